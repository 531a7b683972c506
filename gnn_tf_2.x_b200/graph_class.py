# coding=utf-8
"""Graph containers behind the reference's ``GNN/graph_class.py`` API.

``GraphObject`` is the host-side (NumPy/SciPy) container: same constructor, attributes
and methods as the reference class (``GNN/graph_class.py:14-327``).  ``GraphTensor`` is the
device-side mirror (``GNN/graph_class.py:330-372``): instead of ``tf.SparseTensor`` it holds
torch CUDA tensors plus the destination-sorted CSR / source-sorted CSR^T that the sm_100a
kernels consume.  The CSR is built on the GPU by the C-ABI call ``gnn_csr_build``; its row
order is the order ``tf.sparse.reorder`` gives the transposed COO (``graph_class.py:364-372``).

Differences that are deliberate:
  * NodeGraph of a merged batch is kept as (graph id, coefficient) per node and only
    materialised as the dense (N, G) block-diagonal matrix on attribute access, because the
    dense form (``graph_class.py:313-315``) is impossible at 200k graphs.
  * integer arc endpoints are kept next to the float32 ``arcs`` matrix, so that re-scaling
    ``arcs`` in place (``GNN_utils.py:230,234``) cannot corrupt the structure.
"""
from __future__ import annotations

import os
import shutil
from typing import Optional

import numpy as np
from scipy.sparse import coo_matrix

_FLOATX = 'float32'
_AGGREGATIONS = ('average', 'normalized', 'sum')
# dense NodeGraph is only materialised below this many elements (N * G)
_DENSE_NODEGRAPH_LIMIT = 1 << 28


#######################################################################################################################
## GRAPH OBJECT CLASS #################################################################################################
#######################################################################################################################
class GraphObject:
    """ Host-side graph container. API of the reference ``GraphObject`` (graph_class.py:14). """

    ## CONSTRUCTORS METHODS ###########################################################################################
    def __init__(self, arcs, nodes, targets,
                 problem_based: str = 'n',
                 set_mask=None,
                 output_mask=None,
                 sample_weights=1,
                 NodeGraph=None,
                 ArcNode=None,
                 aggregation_mode: str = 'average',
                 *, _endpoints=None):
        """ CONSTRUCTOR METHOD (graph_class.py:16-77)

        :param arcs: Ordered Arcs Matrix where arcs[i] = [ID Node From | ID NodeTo | Arc Label].
        :param nodes: Ordered Nodes Matrix where nodes[i] = [Node Label].
        :param targets: Targets Array with shape (Num of targeted example [nodes or arcs], dim_target example).
        :param problem_based: (str) 'a' arcs-based, 'g' graph-based, 'n' node-based.
        :param set_mask: Array of {0,1} to define arcs/nodes belonging to a set, when dataset == single GraphObject.
        :param output_mask: Array of {0,1} to define the sub-set of arcs/nodes whose target is known.
        :param sample_weights: target sample weight for loss computation: scalar or array of len(targets).
        :param NodeGraph: Matrix (nodes.shape[0], {Num graphs or 1}) used only when problem_based=='g'.
        :param ArcNode: Matrix of shape (num_of_arcs, num_of_nodes) s.t. A[i,j]=value if arc[i,2]==node[j].
        :param aggregation_mode: (str) 'average' | 'normalized' | 'sum', see buildArcNode.
        """
        self.dtype = _FLOATX
        arcs, nodes, targets = np.asarray(arcs), np.asarray(nodes), np.asarray(targets)

        # integer endpoints are taken BEFORE the float32 cast (exact beyond 2**24 as well);
        # merge() hands them over directly through the private keyword
        if _endpoints is None:
            _endpoints = (np.rint(arcs[:, 0]).astype(np.int64), np.rint(arcs[:, 1]).astype(np.int64))
        self._src, self._dst = (np.ascontiguousarray(i, dtype=np.int64) for i in _endpoints)

        self.arcs = arcs.astype(self.dtype)
        self.nodes = nodes.astype(self.dtype)
        self.targets = targets.astype(self.dtype)
        self.sample_weights = sample_weights * np.ones(self.targets.shape[0])

        self.DIM_NODE_LABEL = nodes.shape[1]
        self.DIM_ARC_LABEL = arcs.shape[1] - 2
        self.DIM_TARGET = targets.shape[1]

        self.problem_based = problem_based
        mask_len = {'n': nodes.shape[0], 'a': arcs.shape[0], 'g': nodes.shape[0]}[problem_based]
        self.set_mask = np.ones(mask_len, dtype=bool) if set_mask is None else np.asarray(set_mask).astype(bool)
        self.output_mask = np.ones(len(self.set_mask), dtype=bool) if output_mask is None \
            else np.asarray(output_mask).astype(bool)
        if len(self.set_mask) != len(self.output_mask):
            raise ValueError('Error - len(<set_mask>) != len(<output_mask>)')

        if aggregation_mode not in _AGGREGATIONS: raise ValueError("ERROR: Unknown aggregation mode")
        self.aggregation_mode = aggregation_mode

        self.ArcNode = self.buildArcNode() if ArcNode is None else coo_matrix(ArcNode).astype(self.dtype)
        self.Adjacency = self.buildAdiacency()
        self._mark_default_structure(ArcNode is None)

        # NodeGraph: (graph id, coefficient) per node + optional user-supplied dense matrix
        self._ng_ids: Optional[np.ndarray] = None
        self._ng_coeff: Optional[np.ndarray] = None
        self._ng_cols: int = 0
        self._ng_dense: Optional[np.ndarray] = None
        if NodeGraph is None:
            built = self.buildNodeGraph(problem_based)
            if built is not None: self._set_nodegraph(built)
        else:
            self._set_nodegraph(NodeGraph)

    # -----------------------------------------------------------------------------------------------------------------
    def copy(self):
        """ Deep copy. As in the reference (graph_class.py:80-87) problem_based and ArcNode are NOT forwarded. """
        return GraphObject(arcs=self.getArcs(), nodes=self.getNodes(), targets=self.getTargets(),
                           set_mask=self.getSetMask(), output_mask=self.getOutputMask(),
                           sample_weights=self.getSampleWeights(), NodeGraph=self._nodegraph_payload(),
                           aggregation_mode=self.aggregation_mode)

    def copy_with_problem(self, problem_based: Optional[str] = None):
        """ deep copy that keeps problem_based and the exact integer endpoints (not in the reference API) """
        return GraphObject(arcs=self.getArcs(), nodes=self.getNodes(), targets=self.getTargets(),
                           problem_based=self.problem_based if problem_based is None else problem_based,
                           set_mask=self.getSetMask(), output_mask=self.getOutputMask(),
                           sample_weights=self.getSampleWeights(), NodeGraph=self._nodegraph_payload(),
                           ArcNode=self.getArcNode(), aggregation_mode=self.aggregation_mode,
                           _endpoints=(self._src.copy(), self._dst.copy()))

    # -----------------------------------------------------------------------------------------------------------------
    def _structure_refs(self):
        a, d = self.ArcNode, self.Adjacency
        return (a, d, a.row, a.col, a.data, d.row, d.col, d.data)

    def _mark_default_structure(self, default: bool) -> None:
        """ remember that ArcNode / Adjacency are exactly what buildArcNode / buildAdiacency made of the endpoints (and
        which objects they are): GraphTensor.fromGraphObject then copies the endpoints once instead of four index arrays
        and knows that the rows of both transposed matrices hold one repeated value """
        self._struct_refs = self._structure_refs() if default else None

    def has_default_structure(self) -> bool:
        refs = getattr(self, '_struct_refs', None)
        return refs is not None and all(x is y for x, y in zip(refs, self._structure_refs()))

    def pin_host_buffers(self) -> None:
        """ move the arrays that GraphTensor.fromGraphObject copies to the device into page-locked host memory
        (numpy views over pinned torch storage), so that the host->device copies run at full PCIe/NVLink-C2C speed """
        import torch
        if not torch.cuda.is_available(): return
        self._pinned = getattr(self, '_pinned', dict())

        def pin(name, arr):
            t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
            self._pinned[name] = t
            return t.numpy()

        default = self.has_default_structure()
        self.arcs, self.nodes, self.targets = pin('arcs', self.arcs), pin('nodes', self.nodes), pin('targets', self.targets)
        self._src32, self._dst32 = pin('src32', self._src.astype(np.int32)), pin('dst32', self._dst.astype(np.int32))
        self._arc_labels, self._arc_labels_of = pin('arc_labels', self.arcs[:, 2:]), self.arcs   # valid while self.arcs is this array
        for name in ('Adjacency', 'ArcNode'):
            m = getattr(self, name)
            m.row, m.col, m.data = pin(name + '.row', m.row.astype(np.int32)), pin(name + '.col', m.col.astype(np.int32)), \
                pin(name + '.data', m.data)
        if default: self._mark_default_structure(True)      # same matrices, new (pinned) arrays

    def host_bytes(self) -> int:
        """ bytes GraphTensor.fromGraphObject copies host->device for this graph """
        total = self.nodes.nbytes + self.targets.nbytes + self.set_mask.nbytes + self.output_mask.nbytes
        total += 4 * self.sample_weights.shape[0]
        # default structure: endpoints once (int32), the two value arrays, arc labels only (the id columns are rebuilt on the device)
        if self.has_default_structure():
            total += 4 * (2 * len(self._src) + len(self.Adjacency.data) + len(self.ArcNode.data)) + 4 * self.arcs.shape[0] * (self.arcs.shape[1] - 2)
        else:
            total += self.arcs.nbytes
            for m in (self.Adjacency, self.ArcNode): total += 4 * (len(m.row) + len(m.col) + len(m.data))
        return int(total)

    ## STRUCTURE BUILDERS #############################################################################################
    def buildArcNode(self):
        """ ArcNode COO of shape (E, N): entry (arc i, dst(i)) = aggregation weight (graph_class.py:98-121).
        'sum' -> 1; 'normalized' -> 1/E (number of ARCS, as the reference code does); 'average' -> 1/indegree(dst). """
        n_arcs, n_nodes = self.arcs.shape[0], self.nodes.shape[0]
        weights = np.ones(n_arcs)
        if self.aggregation_mode == 'normalized':
            weights = weights * float(1 / n_arcs) if n_arcs else weights
        elif self.aggregation_mode == 'average':
            indegree = np.bincount(self._dst, minlength=n_nodes)
            weights = weights / indegree[self._dst]
        return coo_matrix((weights, (np.arange(n_arcs), self._dst)), shape=(n_arcs, n_nodes), dtype=self.dtype)

    # -----------------------------------------------------------------------------------------------------------------
    def buildAdiacency(self):
        """ 'Aggregated' adjacency COO (N, N): entry (src, dst) = ArcNode value of that arc, in arc order;
        duplicate arcs stay duplicate (graph_class.py:90-95). """
        n_nodes = self.nodes.shape[0]
        return coo_matrix((self.ArcNode.data.copy(), (self._src, self._dst)), shape=(n_nodes, n_nodes), dtype=self.dtype)

    # -----------------------------------------------------------------------------------------------------------------
    def setAggregation(self, aggregation_mode: str):
        """ Re-build ArcNode and Adjacency for a new aggregation mode (graph_class.py:124-129). """
        if aggregation_mode not in _AGGREGATIONS: raise ValueError("ERROR: Unknown aggregation mode")
        self.aggregation_mode = aggregation_mode
        self.ArcNode = self.buildArcNode()
        self.Adjacency = self.buildAdiacency()
        self._mark_default_structure(True)

    # -----------------------------------------------------------------------------------------------------------------
    def buildNodeGraph(self, problem_based: str):
        """ (N, 1) matrix of 1/N for graph-based problems, None otherwise (graph_class.py:132-144). """
        if problem_based != 'g': return None
        n_nodes = self.nodes.shape[0]
        return np.ones((n_nodes, 1), dtype=np.float32) * 1 / n_nodes

    ## NODEGRAPH STORAGE ##############################################################################################
    def _set_nodegraph(self, value) -> None:
        """ accept a dense (N, G) matrix or a ('segments', ids, coeff, G) payload """
        if isinstance(value, tuple) and len(value) == 4 and value[0] == 'segments':
            _, ids, coeff, cols = value
            self._ng_ids, self._ng_coeff, self._ng_cols = ids.astype(np.int64), coeff.astype(np.float32), int(cols)
            self._ng_dense = None
            return
        dense = np.asarray(value).astype(self.dtype)
        if dense.ndim == 1: dense = dense[:, None]
        self._ng_cols = dense.shape[1]
        nnz_per_row = (dense != 0).sum(axis=1)
        if dense.shape[0] == self.nodes.shape[0] and np.all(nnz_per_row == 1):
            self._ng_ids = np.argmax(dense != 0, axis=1).astype(np.int64)
            self._ng_coeff = dense[np.arange(dense.shape[0]), self._ng_ids].astype(np.float32)
            self._ng_dense = None
        else:
            # general matrix: kept dense, no segment form
            self._ng_ids, self._ng_coeff, self._ng_dense = None, None, dense

    def _nodegraph_payload(self):
        """ what copy()/merge() forward instead of a dense matrix """
        if self._ng_dense is not None: return self._ng_dense.copy()
        if self._ng_ids is None: return None
        return ('segments', self._ng_ids.copy(), self._ng_coeff.copy(), self._ng_cols)

    @property
    def NodeGraph(self):
        """ dense (N, G) matrix as in the reference; built on access from the segment form """
        if self._ng_dense is not None: return self._ng_dense
        if self._ng_ids is None: return None
        n_nodes = len(self._ng_ids)
        if n_nodes * self._ng_cols > _DENSE_NODEGRAPH_LIMIT:
            raise MemoryError(f'dense NodeGraph ({n_nodes} x {self._ng_cols}) not materialised; use nodegraph_segments()')
        dense = np.zeros((n_nodes, self._ng_cols), dtype=np.float32)
        dense[np.arange(n_nodes), self._ng_ids] = self._ng_coeff
        return dense

    @NodeGraph.setter
    def NodeGraph(self, value):
        if value is None:
            self._ng_ids = self._ng_coeff = self._ng_dense = None
            self._ng_cols = 0
        else:
            self._set_nodegraph(value)

    def nodegraph_segments(self):
        """ (graph id per node [N] int64, coefficient per node [N] float32, number of graphs) or None """
        if self._ng_ids is None: return None
        return self._ng_ids, self._ng_coeff, self._ng_cols

    def has_nodegraph(self) -> bool:
        return self._ng_ids is not None or self._ng_dense is not None

    ## SAVE METHODS ###################################################################################################
    def save(self, graph_folder_path: str) -> None:
        """ save graph in folder as .npy files (graph_class.py:147-152) """
        GraphObject.save_graph(graph_folder_path, self)

    def savetxt(self, graph_folder_path: str, format: str = '%.10g') -> None:
        """ save graph in folder as .txt files (graph_class.py:155-160) """
        GraphObject.save_txt(graph_folder_path, self, format)

    ## GETTERS ########################################################################################################
    def getArcs(self): return self.arcs.copy()

    def getNodes(self): return self.nodes.copy()

    def getTargets(self): return self.targets.copy()

    def getSetMask(self): return self.set_mask.copy()

    def getOutputMask(self): return self.output_mask.copy()

    def getAdjacency(self): return self.Adjacency.copy()

    def getArcNode(self): return self.ArcNode.copy()

    def getNodeGraph(self):
        ng = self.NodeGraph
        return None if ng is None else ng.copy()

    def getSampleWeights(self): return self.sample_weights.copy()

    ## CLASS METHODs ##################################################################################################
    @staticmethod
    def _fresh_folder(path: str) -> str:
        if path[-1] != '/': path += '/'
        if os.path.exists(path): shutil.rmtree(path)
        os.makedirs(path)
        return path

    @staticmethod
    def _optional_items(g) -> dict:
        """ attributes written only when they differ from the constructor defaults (graph_class.py:202-205) """
        items = dict()
        if not all(g.set_mask): items['set_mask'] = g.set_mask
        if not all(g.output_mask): items['output_mask'] = g.output_mask
        if np.any(g.sample_weights != 1): items['sample_weights'] = g.sample_weights
        if g.has_nodegraph() and g.targets.shape[0] > 1: items['NodeGraph'] = g.NodeGraph
        return items

    @classmethod
    def save_graph(cls, graph_folder_path: str, g) -> None:
        """ one .npy per attribute; the folder is re-made (graph_class.py:191-205) """
        folder = cls._fresh_folder(graph_folder_path)
        for name, value in {'arcs': g.arcs, 'nodes': g.nodes, 'targets': g.targets, **cls._optional_items(g)}.items():
            np.save(f'{folder}{name}.npy', value)

    @classmethod
    def save_txt(cls, graph_folder_path: str, g, format: str = '%.10g') -> None:
        """ one .txt per attribute; the folder is re-made (graph_class.py:209-231) """
        folder = cls._fresh_folder(graph_folder_path)
        for name, value in {'arcs': g.arcs, 'nodes': g.nodes, 'targets': g.targets, **cls._optional_items(g)}.items():
            np.savetxt(f'{folder}{name}.txt', value, fmt=format)

    @classmethod
    def _load_folder(cls, graph_folder_path: str, reader, problem_based: str, aggregation_mode: str):
        if graph_folder_path[-1] != '/': graph_folder_path += '/'
        params = {name.rsplit('.')[0]: reader(graph_folder_path + name) for name in os.listdir(graph_folder_path)}
        return cls(**params, problem_based=problem_based, aggregation_mode=aggregation_mode)

    @classmethod
    def load(cls, graph_folder_path: str, problem_based: str, aggregation_mode: str):
        """ load a graph from a folder of .npy files named after constructor arguments (graph_class.py:235-256) """
        return cls._load_folder(graph_folder_path, np.load, problem_based, aggregation_mode)

    @classmethod
    def load_txt(cls, graph_folder_path: str, problem_based: str, aggregation_mode: str):
        """ load a graph from a folder of .txt files (graph_class.py:260-281) """
        return cls._load_folder(graph_folder_path, lambda f: np.loadtxt(f, ndmin=2), problem_based, aggregation_mode)

    # -----------------------------------------------------------------------------------------------------------------
    @classmethod
    def merge(cls, glist, problem_based: str, aggregation_mode: str):
        """ disjoint union of a list of graphs (graph_class.py:284-319): node ids of graph i are shifted by the number
        of nodes before it; for 'g' problems NodeGraph becomes block-diagonal (kept in segment form here). """
        if not (type(glist) == list and all(isinstance(x, (GraphObject, str)) for x in glist)):
            raise TypeError('type of param <glist> must be list of str \'path-like\' or GraphObjects')

        sizes = np.array([g.nodes.shape[0] for g in glist], dtype=np.int64)
        offsets = np.concatenate([[0], np.cumsum(sizes)[:-1]])
        arcs = [g.getArcs() for g in glist]
        for block, off in zip(arcs, offsets): block[:, :2] += off
        arcs = np.concatenate(arcs, axis=0)
        src = np.concatenate([g._src + off for g, off in zip(glist, offsets)])
        dst = np.concatenate([g._dst + off for g, off in zip(glist, offsets)])

        nodegraph = None
        if problem_based == 'g':
            if all(g._ng_ids is not None for g in glist):
                col_off = np.concatenate([[0], np.cumsum([g._ng_cols for g in glist])[:-1]])
                nodegraph = ('segments', np.concatenate([g._ng_ids + o for g, o in zip(glist, col_off)]),
                             np.concatenate([g._ng_coeff for g in glist]), int(sum(g._ng_cols for g in glist)))
            else:
                from scipy.linalg import block_diag
                nodegraph = block_diag(*[g.NodeGraph for g in glist])

        # float32 ids inside `arcs` are exact only below 2**24: hand the exact integer endpoints to the constructor
        return cls(arcs=arcs, nodes=np.concatenate([g.nodes for g in glist], axis=0),
                   targets=np.concatenate([g.targets for g in glist], axis=0), problem_based=problem_based,
                   set_mask=np.concatenate([g.set_mask for g in glist], axis=0),
                   output_mask=np.concatenate([g.output_mask for g in glist], axis=0),
                   sample_weights=np.concatenate([g.sample_weights for g in glist], axis=0),
                   NodeGraph=nodegraph, aggregation_mode=aggregation_mode, _endpoints=(src, dst))

    # -----------------------------------------------------------------------------------------------------------------
    @classmethod
    def fromGraphTensor(cls, g, problem_based: str):
        """ back-conversion (graph_class.py:321-327) """
        nodegraph = None
        if problem_based == 'g': nodegraph = g.nodegraph_payload()
        to_np = lambda t: t.detach().cpu().numpy()
        return cls(arcs=to_np(g.arcs), nodes=to_np(g.nodes), targets=to_np(g.targets), set_mask=to_np(g.set_mask),
                   output_mask=to_np(g.output_mask), sample_weights=to_np(g.sample_weights), NodeGraph=nodegraph,
                   aggregation_mode=g.aggregation_mode, problem_based=problem_based)


#######################################################################################################################
## SPARSE (CSR) VIEW ##################################################################################################
#######################################################################################################################
class SparseCSR:
    """ Row-major sparse matrix on the device: what ``tf.sparse.reorder`` of the transposed COO holds
    (graph_class.py:364-372), stored as CSR.  ``indices``/``values``/``dense_shape`` mirror tf.SparseTensor. """

    def __init__(self, rowptr, col, values, perm, shape, rowptr_T=None, col_T=None, perm_T=None, values_T=None,
                 row_scale=None):
        self.rowptr, self.col, self.values, self.perm = rowptr, col, values, perm
        self.dense_shape = tuple(int(i) for i in shape)
        self.rowptr_T, self.col_T, self.perm_T, self.values_T = rowptr_T, col_T, perm_T, values_T
        # row_scale[n] = common value of row n when every row is uniform ('sum','average','normalized'), else None
        self.row_scale = row_scale

    @property
    def shape(self): return self.dense_shape

    @property
    def nnz(self) -> int: return int(self.col.shape[0])

    @property
    def max_block16_arcs(self) -> int:
        """ largest number of stored entries in an aligned block of 16 rows (computed once, one device read): sizes the
        landing-ring slots of the forward kernel (include/gnn_b200.h, gnn_graph.max_block16_arcs) """
        if getattr(self, '_max_block16', None) is None:
            import torch
            n = int(self.rowptr.shape[0]) - 1
            if n <= 0: self._max_block16 = 0
            else:
                idx = torch.arange(0, n + 16, 16, device=self.rowptr.device).clamp_(max=n)
                b = self.rowptr[idx].to(torch.int64)
                self._max_block16 = int((b[1:] - b[:-1]).max().item())
        return self._max_block16

    @property
    def indices(self):
        """ (nnz, 2) int64 [row, col] pairs in stored (row-major) order """
        import torch
        counts = (self.rowptr[1:] - self.rowptr[:-1]).to(torch.int64)
        rows = torch.repeat_interleave(torch.arange(self.dense_shape[0], device=self.col.device), counts)
        return torch.stack([rows, self.col.to(torch.int64)], dim=1)


#######################################################################################################################
## GRAPH TENSOR CLASS #################################################################################################
#######################################################################################################################
class GraphTensor:
    """ Device-side graph (graph_class.py:330-372). Adjacency and ArcNode are ALREADY transposed: rows = destination. """

    def __init__(self, nodes, arcs, targets, set_mask, output_mask, sample_weights, Adjacency, ArcNode, NodeGraph,
                 aggregation_mode, *, device=None):
        import torch
        from . import _native
        device = _native.default_device() if device is None else torch.device(device)
        f32 = lambda a: torch.as_tensor(a, dtype=torch.float32, device=device) if not isinstance(a, torch.Tensor) \
            else a.to(device=device, dtype=torch.float32)
        bln = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.bool, device=device) if not isinstance(a, torch.Tensor) \
            else a.to(device=device, dtype=torch.bool)
        self.device = device
        self.nodes, self.targets = f32(nodes), f32(targets)
        # arcs: the full [E, 2 + AL] matrix, or ('endpoints+labels', src, dst, labels) device tensors -- the id columns are
        # then only materialised when somebody reads .arcs (the loop needs the labels alone)
        if isinstance(arcs, tuple) and arcs[0] == 'endpoints+labels':
            self._arcs, self._arc_parts = None, (arcs[1], arcs[2], arcs[3])
        else:
            self._arcs, self._arc_parts = f32(arcs), None
        self.sample_weights = f32(sample_weights)
        self.set_mask, self.output_mask = bln(set_mask), bln(output_mask)
        self.aggregation_mode = aggregation_mode
        if not isinstance(Adjacency, SparseCSR) or not isinstance(ArcNode, SparseCSR):
            raise TypeError('Adjacency and ArcNode of a GraphTensor must be SparseCSR (already transposed)')
        self.Adjacency, self.ArcNode = Adjacency, ArcNode

        # NodeGraph: segment form (graph id, coefficient, G) and/or dense tensor
        self._ng_ids = self._ng_coeff = self._ng_dense = None
        self._ng_cols = 0
        if NodeGraph is not None: self._set_nodegraph(NodeGraph)

    @property
    def arcs(self):
        import torch
        if self._arcs is None:
            src, dst, labels = self._arc_parts
            self._arcs = torch.cat([src.to(torch.float32)[:, None], dst.to(torch.float32)[:, None], labels], dim=1)
        return self._arcs

    @arcs.setter
    def arcs(self, value):
        self._arcs, self._arc_parts = value, None

    @property
    def arc_labels(self):
        """ arcs[:, 2:] without materialising the id columns """
        return self._arc_parts[2] if self._arcs is None else self._arcs[:, 2:]

    # -----------------------------------------------------------------------------------------------------------------
    def _set_nodegraph(self, value):
        import torch
        if isinstance(value, tuple) and value[0] == 'segments':
            _, ids, coeff, cols = value
            self._ng_ids = torch.as_tensor(ids, device=self.device).to(torch.int32)
            self._ng_coeff = torch.as_tensor(coeff, device=self.device).to(torch.float32)
            self._ng_cols = int(cols)
        else:
            dense = value if isinstance(value, torch.Tensor) else torch.as_tensor(np.asarray(value))
            dense = dense.to(device=self.device, dtype=torch.float32)
            self._ng_cols = int(dense.shape[1])
            if bool(((dense != 0).sum(dim=1) == 1).all()):
                self._ng_ids = torch.argmax((dense != 0).to(torch.int8), dim=1).to(torch.int32)
                self._ng_coeff = dense.gather(1, self._ng_ids.to(torch.int64)[:, None])[:, 0].contiguous()
            else:
                self._ng_dense = dense

    def nodegraph_payload(self):
        if self._ng_dense is not None: return self._ng_dense.detach().cpu().numpy()
        if self._ng_ids is None: return None
        return ('segments', self._ng_ids.cpu().numpy(), self._ng_coeff.cpu().numpy(), self._ng_cols)

    @property
    def NodeGraph(self):
        """ dense (N, G) tensor as in the reference; built on access """
        import torch
        if self._ng_dense is not None: return self._ng_dense
        if self._ng_ids is None: return None
        n_nodes = int(self._ng_ids.shape[0])
        if n_nodes * self._ng_cols > _DENSE_NODEGRAPH_LIMIT:
            raise MemoryError(f'dense NodeGraph ({n_nodes} x {self._ng_cols}) not materialised')
        dense = torch.zeros((n_nodes, self._ng_cols), dtype=torch.float32, device=self.device)
        dense[torch.arange(n_nodes, device=self.device), self._ng_ids.to(torch.int64)] = self._ng_coeff
        return dense

    @NodeGraph.setter
    def NodeGraph(self, value):
        self._ng_ids = self._ng_coeff = self._ng_dense = None
        self._ng_cols = 0
        if value is not None: self._set_nodegraph(value)

    def has_nodegraph(self) -> bool:
        return self._ng_ids is not None or self._ng_dense is not None

    # cached index tensors (boolean_mask without a device->host sync at every step) ------------------------------------
    def _cached(self, name: str, build):
        # keyed on storage address + in-place version counter: an in-place edit or a reassigned mask rebuilds the index
        cache = self.__dict__.setdefault('_index_cache', dict())
        key = (name, self.set_mask.data_ptr(), self.set_mask._version, self.output_mask.data_ptr(), self.output_mask._version)
        if key not in cache:
            for old in [k for k in cache if k[0] == name]: del cache[old]
            cache[key] = build()
        return cache[key]

    def mask_index(self):
        """ int64 positions where set_mask & output_mask (GNN.py:275) """
        import torch
        return self._cached('mask', lambda: torch.nonzero(self.set_mask & self.output_mask, as_tuple=False)[:, 0])

    def filtered_index(self):
        """ positions, among the rows with output_mask, that also belong to set_mask (GNN_BaseClass.py:405-409) """
        import torch
        return self._cached('filtered', lambda: torch.nonzero(self.set_mask[self.output_mask], as_tuple=False)[:, 0])

    def pool_nodes(self, out_nodes):
        """ NodeGraph^T @ out_nodes (GNN.py:331-332) without the dense matrix: per-graph weighted segment sum.
        Merged batches keep the nodes of a graph contiguous, so the pooling is a CSR over graphs evaluated by the
        library's SpMM kernel (stored order: deterministic); anything else falls back to index_add """
        import torch
        if self._ng_dense is not None: return self._ng_dense.t() @ out_nodes
        idx = self.mask_index()
        if int(out_nodes.shape[0]) == int(self._ng_ids.shape[0]) and self._pool_csr() is not None:
            from .state_loop import segment_pool
            rowptr, col, ids64 = self._pool_csr()
            return segment_pool(rowptr, col, self._ng_coeff, ids64, out_nodes)
        ids = self._ng_ids.to(torch.int64).index_select(0, idx)
        coeff = self._ng_coeff.index_select(0, idx)
        pooled = torch.zeros((self._ng_cols, out_nodes.shape[1]), dtype=out_nodes.dtype, device=out_nodes.device)
        return pooled.index_add(0, ids, coeff[:, None] * out_nodes)

    def _pool_csr(self):
        """ (rowptr over graphs, node index, int64 graph ids) when the graph ids are non-decreasing, else None """
        import torch
        cache = self.__dict__.setdefault('_pool_cache', dict())
        key = (self._ng_ids.data_ptr(), self._ng_ids._version)
        if key not in cache:
            cache.clear()
            ids = self._ng_ids.to(torch.int64)
            if ids.numel() > 1 and bool((ids[1:] < ids[:-1]).any()):
                cache[key] = None
            else:
                counts = torch.bincount(ids, minlength=self._ng_cols)
                rowptr = torch.zeros(self._ng_cols + 1, dtype=torch.int32, device=ids.device)
                rowptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
                cache[key] = (rowptr, torch.arange(ids.numel(), dtype=torch.int32, device=ids.device), ids)
        return cache[key]

    # -----------------------------------------------------------------------------------------------------------------
    def copy(self):
        """ shallow copy sharing the (immutable) device tensors, as the reference does (graph_class.py:347-351) """
        new = GraphTensor.__new__(GraphTensor)
        new.__dict__.update(self.__dict__)
        # the index cache is shared with the copy: its entries are keyed on the masks' storage and version, so a copy whose masks
        # are replaced or edited rebuilds its own entries -- and the per-layer copies LGNN.Loop makes at every call find the
        # indices already built (no torch.nonzero, i.e. no host sync, inside a captured training step)
        new.__dict__['_index_cache'] = self.__dict__.setdefault('_index_cache', dict())
        return new

    # -----------------------------------------------------------------------------------------------------------------
    @classmethod
    def fromGraphObject(cls, g: GraphObject, *, device=None):
        """ GraphObject -> GraphTensor; both sparse matrices are transposed and row-major ordered on the GPU
        (graph_class.py:354-361). """
        if g.has_default_structure():
            # ArcNode / Adjacency as built from the endpoints: Adjacency^T = (row dst, col src), ArcNode^T = (row dst, col arc id).
            # The endpoints travel once (int32), the arc ids are generated on the device, and both builds skip the
            # uniformity read-back (rows of both matrices hold one repeated value in all three aggregation modes), so they
            # run on the GPU while the host goes on copying
            import torch
            from . import _native
            dev = _native.default_device() if device is None else torch.device(device)
            n_nodes, n_arcs = g.nodes.shape[0], g.arcs.shape[0]
            as_dev = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a, dtype=dt), device=dev)
            dst = as_dev(getattr(g, '_dst32', g._dst), np.int32)
            src = as_dev(getattr(g, '_src32', g._src), np.int32)
            adjacency = _native.csr_build(dst, src, as_dev(g.Adjacency.data, np.float32), n_nodes, n_nodes, with_transpose=True,
                                          assume_uniform=True)
            arcnode = _native.csr_build(dst, torch.arange(n_arcs, dtype=torch.int32, device=dev), as_dev(g.ArcNode.data, np.float32),
                                        n_nodes, n_arcs, with_transpose=False, assume_uniform=True)
            labels = g._arc_labels if getattr(g, '_arc_labels_of', None) is g.arcs else g.arcs[:, 2:]
            return cls(nodes=g.nodes, arcs=('endpoints+labels', src, dst, as_dev(labels, np.float32)), targets=g.targets, set_mask=g.set_mask,
                       output_mask=g.output_mask, sample_weights=g.sample_weights, NodeGraph=g._nodegraph_payload(), Adjacency=adjacency,
                       ArcNode=arcnode, aggregation_mode=g.aggregation_mode, device=device)
        else:
            adjacency = cls.COO2SparseTransposedTensor(g.Adjacency, device=device, with_transpose=True)
            arcnode = cls.COO2SparseTransposedTensor(g.ArcNode, device=device, with_transpose=False)
        return cls(nodes=g.nodes, arcs=g.arcs, targets=g.targets, set_mask=g.set_mask, output_mask=g.output_mask,
                   sample_weights=g.sample_weights, NodeGraph=g._nodegraph_payload(), Adjacency=adjacency,
                   ArcNode=arcnode, aggregation_mode=g.aggregation_mode, device=device)

    # -----------------------------------------------------------------------------------------------------------------
    @staticmethod
    def COO2SparseTransposedTensor(coo, *, device=None, with_transpose: bool = False) -> SparseCSR:
        """ transposed, row-major ordered sparse matrix from a scipy COO matrix (graph_class.py:364-372).
        Stored entry order = sort by (coo.col, coo.row), ties in original order. """
        from . import _native
        return _native.csr_from_coo_transposed(coo, device=device, with_transpose=with_transpose)
