# coding=utf-8
"""Metrics dictionary with the reference's keys (``GNN/GNN_metrics.py:152-155``): thin scikit-learn wrappers evaluated
on the host after the outputs have been brought back from the device.  Not on the GPU hot path."""
from __future__ import annotations

import numpy as np


def _np(a):
    return a.detach().cpu().numpy() if hasattr(a, 'detach') else np.asarray(a)


def _confusion_rates(y_true, y_pred, pos_label=0):
    y_true, y_pred = _np(y_true), _np(y_pred)
    pos_t, pos_p = y_true == pos_label, y_pred == pos_label
    tp, fn = np.sum(pos_t & pos_p), np.sum(pos_t & ~pos_p)
    fp, tn = np.sum(~pos_t & pos_p), np.sum(~pos_t & ~pos_p)
    div = lambda a, b: float(a) / float(b) if b else 0.0
    return dict(Tpr=div(tp, tp + fn), Tnr=div(tn, tn + fp), Fpr=div(fp, fp + tn), Fnr=div(fn, fn + tp))


def TPR(y_true, y_pred, pos_label=0, **_): return _confusion_rates(y_true, y_pred, pos_label)['Tpr']
def TNR(y_true, y_pred, pos_label=0, **_): return _confusion_rates(y_true, y_pred, pos_label)['Tnr']
def FPR(y_true, y_pred, pos_label=0, **_): return _confusion_rates(y_true, y_pred, pos_label)['Fpr']
def FNR(y_true, y_pred, pos_label=0, **_): return _confusion_rates(y_true, y_pred, pos_label)['Fnr']


def _sk(name):
    def metric(y_true, y_pred, **kwargs):
        import sklearn.metrics as skm
        return getattr(skm, name)(_np(y_true), _np(y_pred), **kwargs)
    metric.__name__ = name
    return metric


Metrics = {'Acc': _sk('accuracy_score'), 'Bacc': _sk('balanced_accuracy_score'), 'Js': _sk('jaccard_score'),
           'Ck': _sk('cohen_kappa_score'), 'Prec': _sk('precision_score'), 'Rec': _sk('recall_score'),
           'Fs': _sk('f1_score'), 'Tpr': TPR, 'Tnr': TNR, 'Fpr': FPR, 'Fnr': FNR}


def ROC(targets, y_score, outdir: str, micro_and_macro: bool = False, pos_label=0):
    """ ROC data per class (the reference also plots with matplotlib, GNN_metrics.py:108-138; plotting is skipped when
    matplotlib is not installed). Returns {class: (fpr, tpr, auc)} """
    import sklearn.metrics as skm
    targets, y_score = _np(targets), _np(y_score)
    curves = dict()
    for c in range(targets.shape[1]):
        fpr, tpr, _ = skm.roc_curve(targets[:, c], y_score[:, c])
        curves[c] = (fpr, tpr, skm.auc(fpr, tpr))
    if micro_and_macro:
        fpr, tpr, _ = skm.roc_curve(targets.ravel(), y_score.ravel())
        curves['micro'] = (fpr, tpr, skm.auc(fpr, tpr))
    try:
        import matplotlib
        matplotlib.use('Agg')
        import matplotlib.pyplot as plt
        import os
        os.makedirs(outdir, exist_ok=True)
        for name, (fpr, tpr, auc) in curves.items(): plt.plot(fpr, tpr, label=f'{name} (AUC {auc:.3f})')
        plt.legend(), plt.xlabel('FPR'), plt.ylabel('TPR')
        plt.savefig(os.path.join(outdir, 'ROC.png'))
        plt.close()
    except ImportError:
        pass
    return curves


def PRISOFS(targets, y_score, outdir: str, pos_label=0):
    """ precision-recall data per class (GNN_metrics.py:142-148) """
    import sklearn.metrics as skm
    targets, y_score = _np(targets), _np(y_score)
    return {c: skm.precision_recall_curve(targets[:, c], y_score[:, c])[:2] for c in range(targets.shape[1])}
