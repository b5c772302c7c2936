"""Drop-in mirror of the counting half of the reference's ``metrics.py`` on the GPU.

``er_score`` / ``f1_score`` / ``cos_sim`` keep the reference's Keras-metric calling
convention ``(y_true, y_pred) -> [B]`` or scalar (metrics.py:217-298).  The integer core --
``(n_true, n_pred, correct)`` per sample and ``(TP, FP, FN)`` -- is one integer reduction
kernel (``iris_metric_counts``); for multi-GPU runs the count vector is what
``challenge_b200.dist`` all-reduces.
"""
import numpy as np

from . import _ops as O
from .dist import f1_from_counts
from .engine import get_engine


def _first(t):
    return t[0] if isinstance(t, tuple) else t


def er_score(threshold=0.5, smoothing=True):
    """metrics.py:217-274.  ``smoothing=True`` average-pools ``y_pred`` with
    ``AveragePooling1D(31, padding='same')`` first (strides default to the pool size, so the
    predictions are scored on the pooled time base, exactly as the reference does)."""
    def er(y_true, y_pred):
        eng = get_engine()
        yt, yp = O.dev(y_true), O.dev(y_pred)
        if smoothing:
            k = int(0.5 * 16000) // 256
            B, T, K = (int(s) for s in yp.shape)
            pooled = O.empty((B, -(-T // k), K))
            O.call('iris_op_avg_pool_time', O.ptr(yp), O.ptr(pooled), B, T, K, k, 0)
            yp = pooled
        if yt.shape != yp.shape:
            # the reference compares frame indices of different time bases in this case
            # (metrics.py:259-266); the counting kernel takes one frame count
            raise NotImplementedError('er_score(smoothing=True) with T > 31 scores y_true and the '
                                      'pooled y_pred on different time bases; use smoothing=False '
                                      '(sj_train.py:457)')
        _, _, er_ = eng.metric_counts(yt, yp, threshold=float(threshold))
        return er_
    return er


def er_counts(y_true, y_pred, threshold=0.5):
    """The integer core of ``er_score(smoothing=False)``: int32 ``[B, 3]`` rows
    ``(n_true, n_pred, correct)`` (metrics.py:229-266)."""
    triples, _, _ = get_engine().metric_counts(O.dev(y_true), O.dev(y_pred),
                                               threshold=float(threshold), want_er=False)
    return triples


def cos_sim(y_true, y_pred):
    """metrics.py:277-287."""
    yt, yp = O.dev(_first(y_true)), O.dev(_first(y_pred))
    B, T, K = (int(s) for s in yt.shape)
    out = O.empty((B,))
    O.call('iris_op_cos_sim', O.ptr(yt), O.ptr(yp), O.ptr(out), B, T, K)
    return out


def f1_score():
    """metrics.py:290-298 -- ``tfa.metrics.F1Score(3, threshold=0.5, average='micro')`` held in
    a closure: the TP / FP / FN counts ACCUMULATE over every call and are never reset."""
    state = {}

    def f1_score(y_true, y_pred):
        import torch
        eng = get_engine()
        yt, yp = O.dev(_first(y_true)), O.dev(_first(y_pred))
        if 'tpfpfn' not in state:
            state['tpfpfn'] = torch.zeros(3, dtype=torch.int64, device=eng.device)
        eng.metric_counts(yt, yp, threshold=0.5, tpfpfn=state['tpfpfn'], want_er=False)
        tp, fp, fn = (int(v) for v in state['tpfpfn'].cpu().numpy())
        return np.float32(f1_from_counts(tp, fp, fn))
    f1_score.state = state
    return f1_score
