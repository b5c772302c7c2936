"""Oracle restatement of the evaluation-side chain of the reference's ``metrics.py``
(``evaluate`` 30-90, ``Challenge_Metric.get_start_end_frame`` 109-133, ``get_er`` 176-193,
``output_to_metric`` 196-214).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference holds no test or fixture for these functions (the only data
is ``sample_answer.json``, an INPUT of ``evaluate``); the restatement follows the cited lines
and the TF-2.2 semantics flagged [TF-sem], written as plain loops.  Where TensorFlow leaves an
order unspecified (``tf.argsort`` is not stable; the accumulation order inside
``overlap_and_add``) the choice made here -- stable sort, ascending window order -- is stated
at the line and is the same on the GPU.
"""
import numpy as np

f32 = np.float32


def frame_windows(x, frame_length, frame_step):
    """metrics.py:60-61 -- ``tf.signal.frame(x, n_frame, hop, pad_end=True, axis=-2)`` then
    ``transpose (1, 0, 2, 3)``.  x: [mel, T, C] -> [n_win, mel, n_frame, C].
    [TF-sem] with ``pad_end`` the number of frames is ``ceil(T / frame_step)`` and frames that
    run past the end are zero-padded."""
    x = np.asarray(x, f32)
    M, T, C = x.shape
    n_win = -(-T // frame_step)
    out = np.zeros((n_win, M, frame_length, C), f32)
    for w in range(n_win):
        lo = w * frame_step
        hi = min(lo + frame_length, T)
        if hi > lo:
            out[w, :, :hi - lo, :] = x[:, lo:hi, :]
    return out


def overlap_average(preds, frame_step, frame_len, upsample=1):
    """metrics.py:67-75 -- ``UpSampling1D(r)`` (each step repeated r times), transpose to
    [K, n_win, n_frame], ``overlap_and_add`` of the predictions and of ones, ``[..., :frame_len]``,
    divide, transpose back.  preds: [n_win, n_p, K] -> [frame_len, K].
    [TF-sem] ``overlap_and_add`` output length is ``(n_win - 1) * step + n_frame``; positions no
    window covers are 0 / 0 = NaN.  Accumulation in ascending window order (unspecified in TF;
    irrelevant for <= 2 overlapping windows)."""
    preds = np.asarray(preds, f32)
    if upsample > 1:
        preds = np.repeat(preds, int(upsample), axis=1)
    n_win, F, K = preds.shape
    total = (n_win - 1) * frame_step + F
    acc = np.zeros((total, K), f32)
    cnt = np.zeros((total, K), f32)
    for w in range(n_win):
        acc[w * frame_step:w * frame_step + F] += preds[w]
        cnt[w * frame_step:w * frame_step + F] += f32(1)
    with np.errstate(invalid='ignore', divide='ignore'):
        out = acc[:frame_len] / cnt[:frame_len]
    if out.shape[0] < frame_len:   # [..., :frame_len] of a shorter signal is the signal
        return out
    return out


def avg_pool_same_stride1(x, k):
    """``AveragePooling1D(k, 1, padding='same')`` on [L, K] (metrics.py:79).  [TF-sem] SAME pads
    ``(k - 1) // 2`` before and the rest after; the mean is over the VALID cells only; the CPU
    kernel accumulates the inputs of a window in ascending order."""
    x = np.asarray(x, f32)
    L, K = x.shape
    before = (k - 1) // 2
    out = np.zeros_like(x)
    for t in range(L):
        lo, hi = max(t - before, 0), min(t - before + k, L)
        s = np.zeros(K, f32)
        for u in range(lo, hi):
            s = (s + x[u]).astype(f32)
        out[t] = s / f32(hi - lo)
    return out


def max_pool_same_stride1(x, k):
    """``MaxPooling1D(k, 1, padding='same')`` on [L, K] (metrics.py:80); padding never wins."""
    x = np.asarray(x, f32)
    L, K = x.shape
    before = (k - 1) // 2
    out = np.zeros_like(x)
    for t in range(L):
        lo, hi = max(t - before, 0), min(t - before + k, L)
        out[t] = x[lo:hi].max(axis=0)
    return out


def smooth(preds, sr=16000, hop=256, threshold=0.5):
    """metrics.py:77-81: kernel = int(0.5 * sr) // hop = 31; avg pool 31, max pool 124, >= 0.5."""
    k = int(0.5 * sr) // hop
    y = avg_pool_same_stride1(preds, k)
    y = max_pool_same_stride1(y, k * 4)
    return (y >= f32(threshold)).astype(f32)


def get_start_end_frame(data):
    """metrics.py:109-133 -- per class the time indices where ``data`` differs from its
    predecessor (a zero row before t = 0), paired up as (start, next change - 1); an odd count is
    closed with ``len(data)``.  Returns three int64 [n, 2] arrays."""
    data = np.asarray(data, f32)
    prev = np.concatenate([np.zeros((1, data.shape[1]), f32), data[:-1]], 0)
    out = []
    for c in range(3):
        idx = np.nonzero(prev[:, c] != data[:, c])[0].astype(np.int64)
        if idx.shape[0] % 2 != 0:
            idx = np.concatenate([idx, np.asarray([len(data)], np.int64)])
        idx = idx.reshape(-1, 2)
        out.append(np.stack([idx[:, 0], idx[:, 1] - 1], 1))
    return tuple(out)


def output_to_metric(hop, sr):
    """metrics.py:196-214 -- rows (class, int32(((start + end) / 2) * hop / sr)); the arithmetic
    is float64 ([TF-sem] int64 / int is true division in float64) and the cast truncates."""
    def _f(cls0, cls1, cls2):
        rows = []
        for c, items in enumerate((cls0, cls1, cls2)):
            for s, e in np.asarray(items, np.int64).reshape(-1, 2):
                v = ((np.float64(s) + np.float64(e)) / 2) * hop / sr
                rows.append([c, int(np.trunc(v))])
        return np.asarray(rows, np.int32).reshape(-1, 2)
    return _f


def get_er(gt, predict):
    """metrics.py:176-193 -- greedy matching.  Predictions sorted by time, ground truth by start
    time (stable sort here; ``tf.argsort`` leaves ties unspecified); every ground-truth row takes
    the first remaining prediction of its class whose time lies in [start, end].
    Returns ``(er, N, answer)`` with ``er = (N - answer) / len(gt)``."""
    predict = np.asarray(predict, np.int64).reshape(-1, 2)
    gt = np.asarray(gt, np.int64).reshape(-1, 3)
    pred = [tuple(r) for r in predict[np.argsort(predict[:, 1], kind='stable')]]
    gts = gt[np.argsort(gt[:, 1], kind='stable')]
    N = len(pred) + len(gts)
    answer = 0
    for g in gts:
        for i, p in enumerate(pred):
            if g[1] <= p[1] <= g[2] and g[0] == p[0]:
                answer += 2
                del pred[i]
                break
    return (N - answer) / len(gts), N, answer


def evaluate_one(features, predict_fn, gt, n_frame, n_chan, overlap_hop=512, upsample=1,
                 sr=16000, hop=256):
    """metrics.py:59-87 for one file, given its model-ready features [mel, T, C]."""
    frame_len = features.shape[-2]
    win = frame_windows(features, n_frame, overlap_hop)
    preds = np.asarray(predict_fn(win[..., :n_chan]), f32)
    merged = overlap_average(preds, overlap_hop, frame_len, upsample)
    y = smooth(merged, sr, hop)
    cls = get_start_end_frame(y)
    rows = output_to_metric(hop, sr)(*cls)
    er, N, answer = get_er(gt, rows)
    return er, dict(windows=win, merged=merged, smoothed=y, cls=cls, rows=rows, N=N, answer=answer)
