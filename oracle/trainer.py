"""Oracle restatement of the data-path functions of the reference's ``trainer.py``
(TEST INFRASTRUCTURE ONLY).  PARITY UNPINNED (the reference holds no test for them); plain
numpy in the cited lines' op order, TF semantics flagged [TF-sem]."""
import numpy as np

f32 = np.float32


def preprocess_labels(multiplier):
    """trainer.py:86-94.  [TF-sem] ``avg_pool1d(y, 2, strides=2, padding='SAME')`` pads on the
    right only (total pad <= 1) and averages over the valid cells."""
    def _f(x, y):
        y = np.asarray(y, f32)
        for _ in range(5):
            B, T, K = y.shape
            n = (T + 1) // 2
            out = np.zeros((B, n, K), f32)
            for i in range(n):
                cells = y[:, 2 * i:min(2 * i + 2, T)]
                out[:, i] = (cells.sum(axis=1, dtype=f32) / f32(cells.shape[1])) * f32(2)
            y = out
        return x, (y * f32(multiplier)).astype(f32)
    return _f


def to_density_labels(x, y):
    """trainer.py:97-104: safe_div by the per-voice total over (frames, classes), sum over voices."""
    y = np.asarray(y, f32)
    tot = y.sum(axis=(-2, -1), keepdims=True, dtype=f32)
    y = y / np.maximum(tot, f32(1e-8))
    out = np.zeros(y.shape[:-3] + y.shape[-2:], f32)
    for v in range(y.shape[-3]):
        out = (out + y[..., v, :, :]).astype(f32)
    return x, out


def minmax_log_on_mel(mel):
    """trainer.py:61-76."""
    mel = np.asarray(mel, f32)
    axis = tuple(range(1, mel.ndim))
    mx = mel.max(axis=axis, keepdims=True)
    mn = mel.min(axis=axis, keepdims=True)
    mel = (mel - mn) / np.maximum(mx - mn, f32(1e-8))
    return np.log(mel + f32(1e-8)).astype(f32)
