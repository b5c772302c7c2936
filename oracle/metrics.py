"""Oracle restatement of the reference's ``metrics.py`` counting (TEST INFRASTRUCTURE ONLY)."""
import numpy as np

from .data_utils import safe_div


def _where(x):
    """``tf.where(x)`` -- coordinates of non-zeros in row-major order, int64 ``[n, ndim]``."""
    return np.argwhere(x).astype(np.int64)


def _sort_rows(idx):
    """metrics.py:235-240.  ``tf.gather(x, order, -1)`` passes ``-1`` as
    ``validate_indices`` so BOTH gathers are row gathers [TF-sem]: rows sorted by
    class (stable), then by batch (stable) => order (b, c, t)."""
    idx = idx[np.argsort(idx[:, -1], kind='stable')]
    idx = idx[np.argsort(idx[:, 0], kind='stable')]
    return idx


def er_parts(y_true, y_pred, threshold=0.5, smoothing=False):
    """metrics.py:220-266 up to ``correct_per_sample``.

    Returns the integer core ``(n_true[B], n_pred[B], correct[B])``.
    """
    f32 = np.float32
    y_true = (np.asarray(y_true, f32) >= f32(threshold)).astype(np.int32)      # 221
    y_pred = np.asarray(y_pred, f32)
    if smoothing:
        from .data_utils import _avg_pool1d_same
        k = int(0.5 * 16000) // 256                                             # 223
        # AveragePooling1D(k, padding='same'): strides default to pool_size
        y_pred = _avg_pool1d_same(y_pred, k)
    y_pred = (y_pred >= f32(threshold)).astype(np.int32)                        # 225
    B = y_pred.shape[0]

    def edges(y):
        starts = np.clip(y - np.pad(y, [(0, 0), (1, 0), (0, 0)])[:, :-1], 0, 1)
        ends = np.clip(y - np.pad(y, [(0, 0), (0, 1), (0, 0)])[:, 1:], 0, 1)
        n = starts.astype(f32).sum(axis=(1, 2))
        return _sort_rows(_where(starts)), _sort_rows(_where(ends)), n

    true_starts, true_ends, n_true = edges(y_true)
    pred_starts, pred_ends, n_pred = edges(y_pred)

    middle = ((pred_starts + pred_ends) / 2).astype(np.int64)                   # 256

    # [N, 2, M] compare of (b, c) then time-window test (259-266)
    correct = (true_starts[:, ::2, None] == middle.T[None, ::2])
    correct = correct.astype(f32).min(axis=1)
    mid_time = middle[:, 1:2].T
    correct = correct * (true_starts[:, 1:2] <= mid_time).astype(f32)
    correct = correct * (true_ends[:, 1:2] >= mid_time).astype(f32)
    correct = np.pad(correct, [(0, 0), (0, 1)]).max(axis=-1)

    onehot = np.zeros((len(true_starts), B), f32)
    onehot[np.arange(len(true_starts)), true_starts[:, 0]] = 1
    correct_per_sample = (onehot * correct[:, None]).sum(axis=0)
    return (n_true.astype(np.int32), n_pred.astype(np.int32),
            correct_per_sample.astype(np.int32))


def er_from_parts(n_true, n_pred, correct):
    """metrics.py:268-273 -- fp32 score per sample."""
    f32 = np.float32
    n_true = np.asarray(n_true, f32)
    score = n_true + np.asarray(n_pred, f32) - f32(2) * np.asarray(correct, f32)
    hi = n_true.max() if n_true.size else f32(0)
    with np.errstate(divide='ignore', invalid='ignore'):
        # tf.clip_by_value(x, lo, hi) = minimum(maximum(x, lo), hi)
        score = score / np.minimum(np.maximum(n_true, f32(1)), hi)
    return score.astype(f32)


def er_score(threshold=0.5, smoothing=True):
    """metrics.py:217-274."""
    def er(y_true, y_pred):
        return er_from_parts(*er_parts(y_true, y_pred, threshold, smoothing))
    return er


def f1_counts(y_true, y_pred, threshold=0.5):
    """tfa.metrics.F1Score(3, threshold=0.5, 'micro').update_state [TFA-sem]
    (constructed at metrics.py:291): ``pred = y_pred > threshold`` (strict),
    TP = sum(pred*true), FP = sum(pred*(1-true)), FN = sum((1-pred)*true)."""
    f32 = np.float32
    y_true = np.asarray(y_true, f32)
    pred = (np.asarray(y_pred, f32) > f32(threshold)).astype(f32)
    tp = np.sum(pred * y_true, dtype=np.float64)
    fp = np.sum(pred * (1 - y_true), dtype=np.float64)
    fn = np.sum((1 - pred) * y_true, dtype=np.float64)
    return int(round(tp)), int(round(fp)), int(round(fn))


def f1_from_counts(tp, fp, fn):
    """tfa FBetaScore.result() with beta=1 [TFA-sem]: divide_no_nan everywhere."""
    f32 = np.float32

    def dnn(a, b):
        return f32(0) if b == 0 else f32(a) / f32(b)

    tp, fp, fn = f32(tp), f32(fp), f32(fn)
    precision = dnn(tp, tp + fp)
    recall = dnn(tp, tp + fn)
    mul = precision * recall
    add = precision + recall
    return f32(dnn(mul, add) * f32(2))


class F1State:
    """``f1_score()`` (metrics.py:290-298) captures ONE stateful Metric object in a
    closure and never resets it, so counts accumulate across calls."""

    def __init__(self):
        self.tp = self.fp = self.fn = 0

    def __call__(self, y_true, y_pred):
        if isinstance(y_true, tuple):
            y_true = y_true[0]
        if isinstance(y_pred, tuple):
            y_pred = y_pred[0]
        tp, fp, fn = f1_counts(y_true, y_pred)
        self.tp += tp
        self.fp += fp
        self.fn += fn
        return f1_from_counts(self.tp, self.fp, self.fn)


def f1_score():
    return F1State()


def cos_sim(y_true, y_pred):
    """metrics.py:277-287 (keras cosine_similarity = -sum(l2norm(a)*l2norm(b)))."""
    f32 = np.float32
    if isinstance(y_true, tuple):
        y_true = y_true[0]
    if isinstance(y_pred, tuple):
        y_pred = y_pred[0]
    y_true = np.asarray(y_true, f32)
    y_pred = np.asarray(y_pred, f32)
    m = (y_true.sum(axis=-2) > 0.).astype(f32)
    m = safe_div(m, m.sum(axis=-1, keepdims=True))

    def l2n(x):
        ss = np.sum(x * x, axis=-2, keepdims=True, dtype=f32)
        return x * (f32(1) / np.sqrt(np.maximum(ss, f32(1e-12))))

    cs = -np.sum(l2n(y_true) * l2n(y_pred), axis=-2, dtype=f32)
    return np.sum(cs * m, axis=-1, dtype=f32)


def get_er(gt, predict):
    """metrics.py:176-193 (eval only): greedy first-match removal."""
    predict = [list(map(int, p)) for p in predict]
    gt = [list(map(int, g)) for g in gt]
    predict = [predict[i] for i in np.argsort([p[1] for p in predict], kind='stable')]
    gt = [gt[i] for i in np.argsort([g[1] for g in gt], kind='stable')]
    N = len(predict) + len(gt)
    answer = 0
    for g in gt:
        hit = None
        for i, p in enumerate(predict):
            if g[1] <= p[1] <= g[2] and g[0] == p[0]:
                answer += 2
                hit = i
                break
        if hit is not None:
            predict = predict[:hit] + predict[hit + 1:]
    return (N - answer) / len(gt)
