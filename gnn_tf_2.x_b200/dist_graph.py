# coding=utf-8
"""Multi-GPU execution of the state loop (placeholder until the partitioned path lands in this round)."""


def bench_partitioned(g_host, wl, build_gnn, args, device, rank, world):
    raise NotImplementedError('node-range partitioned execution is not implemented yet')
