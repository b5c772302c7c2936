"""Drop-in mirror of the counting half of the reference's ``metrics.py`` on the GPU.

``er_score`` / ``f1_score`` / ``cos_sim`` keep the reference's Keras-metric calling
convention ``(y_true, y_pred) -> [B]`` or scalar (metrics.py:217-298).  The integer core --
``(n_true, n_pred, correct)`` per sample and ``(TP, FP, FN)`` -- is one integer reduction
kernel (``iris_metric_counts``); for multi-GPU runs the count vector is what
``challenge_b200.dist`` all-reduces.

The evaluation-side chain of ``metrics.evaluate`` (metrics.py:40-90) is here too: windows over a
file's features, overlap-and-add averaging of the per-window predictions, the smoothing pools,
``Challenge_Metric.get_start_end_frame``, ``output_to_metric`` and ``get_er`` -- every stage
between the model and the score is a kernel of libiris (k_eval.cu), so nothing but the final
three integers leaves the device.
"""
import json
import os
from glob import glob

import numpy as np

from . import _ops as O
from .dist import f1_from_counts
from .engine import get_engine

label_downsample_model = (3, 6, 7, 8, 9)   # utils.py:7


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
            # the reference compares frame indices of the two time bases as they are
            # (metrics.py:256-266); the counting kernel keeps the two frame counts apart
            return eng.er_counts_pooled(yt, yp, threshold=float(threshold))[1]
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


# ---------------------------------------------------------------------------------------------
# evaluation-side chain (metrics.py:30-214)
# ---------------------------------------------------------------------------------------------
def _i32dev(x):
    import torch
    eng = get_engine()
    return torch.as_tensor(np.ascontiguousarray(np.asarray(x, np.int64).astype(np.int32))).to(eng.device)


def frame_windows(inputs, n_frame, overlap_hop=512):
    """metrics.py:60-61 -- ``tf.signal.frame(inputs, n_frame, overlap_hop, pad_end=True, axis=-2)``
    followed by ``transpose (1, 0, 2, 3)``: [mel, T, C] -> [n_win, mel, n_frame, C]."""
    t = O.dev(inputs)
    if t.dim() != 3:
        raise ValueError('frame_windows expects [freq, time, chan]')
    M, T, Cc = (int(s) for s in t.shape)
    n_win = -(-T // int(overlap_hop))
    out = O.empty((n_win, M, int(n_frame), Cc))
    O.call('iris_op_eval_windows', O.ptr(t), O.ptr(out), M, T, Cc, int(n_frame), int(overlap_hop), n_win)
    return out


def overlap_average(preds, overlap_hop, frame_len, upsample=1):
    """metrics.py:67-75 -- ``UpSampling1D`` + both ``overlap_and_add`` calls + the division:
    [n_win, n_p, K] -> [frame_len, K]."""
    p = O.dev(preds)
    n_win, n_p, K = (int(s) for s in p.shape)
    total = (n_win - 1) * int(overlap_hop) + n_p * int(upsample)
    L_ = min(int(frame_len), total)                      # [..., :frame_len]
    out = O.empty((L_, K))
    O.call('iris_op_eval_merge', O.ptr(p), O.ptr(out), n_win, n_p, K, int(upsample), int(overlap_hop), L_)
    return out


def smooth_predictions(preds, sr=16000, hop=256, threshold=0.5):
    """metrics.py:77-81 -- AveragePooling1D(31, 1, 'same'), MaxPooling1D(124, 1, 'same'),
    ``>= 0.5`` as floats.  preds [time, K]."""
    p = O.dev(preds)
    L_, K = (int(s) for s in p.shape)
    k = int(0.5 * sr) // hop
    tmp, out = O.empty((L_, K)), O.empty((L_, K))
    O.call('iris_op_eval_smooth', O.ptr(p), O.ptr(tmp), O.ptr(out), L_, K, k, 4 * k, float(threshold))
    return out


def _events(data, hop, sr):
    """-> (rows int32 [n, 4] = (class, start, end, second), per-class counts) on the device."""
    import torch
    eng = get_engine()
    y = O.dev(data)
    L_, K = (int(s) for s in y.shape)
    max_rows = K * (L_ // 2 + 1)
    rows = torch.empty((max_rows, 4), dtype=torch.int32, device=eng.device)
    n_rows = torch.empty((1 + K,), dtype=torch.int32, device=eng.device)
    O.call('iris_op_eval_events', O.ptr(y), L_, K, int(hop), int(sr), O.ptr(rows), max_rows, O.ptr(n_rows))
    return rows, n_rows


class Challenge_Metric:
    """metrics.py:93-173 (the frame-level half; ``get_second_answer`` calls a method the
    reference does not define and is not mirrored)."""

    def __init__(self, sr=16000, hop=256) -> None:
        self.sr = sr
        self.hop = hop

    def get_start_end_frame(self, data):
        """metrics.py:109-133 -- three int64 [n, 2] tensors of (start, end) frames."""
        rows, n_rows = _events(data, self.hop, self.sr)
        counts = [int(v) for v in n_rows.cpu().numpy()]
        out, base = [], 0
        for c in range(3):
            n = counts[1 + c] if 1 + c < len(counts) else 0
            out.append(rows[base:base + n, 1:3].long())
            base += n
        return tuple(out)

    def get_start_end_time(self, data):
        """metrics.py:101-107 -- frames -> rounded seconds, duplicates dropped (first kept)."""
        import torch
        outs = []
        for d in self.get_start_end_frame(data):
            sec = torch.round(d.double() * self.hop / self.sr).to(torch.int32).cpu().numpy()
            _, first = np.unique(sec, return_index=True, axis=0) if len(sec) else (None, np.zeros(0, np.int64))
            outs.append(torch.as_tensor(sec[first].reshape(-1, 2)))
        return tuple(outs)


def output_to_metric(hop, sr):
    """metrics.py:196-214 -- [n, 2] int32 rows (class, int32(((start + end) / 2) * hop / sr)),
    float64 arithmetic, truncating cast.  (The fused ``evaluate`` below takes these rows straight
    from the event kernel; this form serves callers that hold the per-class lists.)"""
    def output_to_metric_(cls0, cls1, cls2):
        import torch
        rows = []
        for c, items in enumerate((cls0, cls1, cls2)):
            it = torch.as_tensor(items).reshape(-1, 2).to(torch.float64)
            sec = (((it[:, 0] + it[:, 1]) / 2) * hop / sr).to(torch.int32)
            rows.append(torch.stack([torch.full_like(sec, c), sec], 1))
        return torch.cat(rows, 0) if rows else torch.zeros((0, 2), dtype=torch.int32)
    return output_to_metric_


def er_counts_events(gt, predict):
    """The integer core of ``get_er``: int32 device tensor ``(N, answer, len(gt))``."""
    import torch
    eng = get_engine()
    g = _i32dev(np.asarray(gt).reshape(-1, 3))
    if isinstance(predict, tuple):               # (rows [max, 4], n_rows) from the event kernel
        rows, n_rows = predict
        pr, stride, col, n_ptr, n_max = rows, 4, 3, O.ptr(n_rows), int(rows.shape[0])
    else:
        pr = predict.to(device=eng.device, dtype=torch.int32).contiguous() if hasattr(predict, 'to') \
            else _i32dev(np.asarray(predict).reshape(-1, 2))
        stride, col, n_ptr, n_max = 2, 1, None, int(pr.shape[0])
    out = torch.empty((3,), dtype=torch.int32, device=eng.device)
    O.call('iris_op_get_er', O.ptr(g) if g.numel() else None, int(g.shape[0]),
           O.ptr(pr) if n_max else None, stride, col, n_ptr, n_max, O.ptr(out))
    return out


def get_er(gt, predict):
    """metrics.py:176-193 -- ``(N - answer) / len(gt)`` of the greedy matching (ties in the two
    sorts are kept in input order; ``tf.argsort`` leaves them unspecified)."""
    N, answer, m = (int(v) for v in er_counts_events(gt, predict).cpu().numpy())
    return (N - answer) / m


def _predict(model, x):
    fn = getattr(model, 'predict', None) or model
    return fn(x)


def evaluate(config, model, overlap_hop=512, verbose: bool = False, *, wavs=None, answers=None):
    """metrics.py:30-90.  ``wavs`` (name -> waveform [chan, samples] or file name) and ``answers``
    (name -> [[class, start, end], ...]) default to the reference's ``glob('*.wav')`` and
    ``sample_answer.json`` in the working directory.  ``model`` is anything with ``.predict`` (or
    a callable) mapping [n_win, mel, n_frame, n_chan] to [n_win, time, 3]."""
    from .data_utils import (load_wav, mono_chan, stereo_mono, random_merge_aug, stft_filter,
                             minmax, log_on_mel, speech_enhancement_preprocess)
    from .transforms import complex_to_magphase, magphase_to_mel
    if answers is None:
        with open('sample_answer.json') as f:
            answers = json.load(f)['task2_answer']
    if wavs is None:
        wavs = {os.path.basename(p)[:-4]: p for p in sorted(glob('*.wav'))}
    sr, hop = 16000, 256
    final_score = []
    for name in sorted(wavs):
        inputs = load_wav(wavs[name])
        if config.n_chan == 1:
            inputs = mono_chan(inputs)
        elif config.n_chan == 3:
            inputs = stereo_mono(inputs)
        elif config.n_chan > 3:
            inputs = random_merge_aug(config.n_chan)(inputs, None)
        if config.model_type != 'se':
            inputs = stft_filter(int(round(256 * 1000 / 16000)))(inputs)
            inputs = complex_to_magphase(inputs)
            inputs = magphase_to_mel(config.n_mels)(inputs)
            inputs = minmax(inputs)          # unbatched input: per-mel-row, as in the reference
            inputs = log_on_mel(inputs)
        else:
            inputs = speech_enhancement_preprocess(inputs)
        frame_len = int(inputs.shape[-2])
        windows = frame_windows(inputs, config.n_frame, overlap_hop)
        preds = _predict(model, windows[..., :config.n_chan])
        if config.model_type == 'se' and config.v == 9:
            preds = preds[0]
        preds = O.dev(preds)
        up = 1
        if config.v in label_downsample_model:
            up = int(config.n_frame / int(preds.shape[-2]))
        merged = overlap_average(preds, overlap_hop, frame_len, up)
        y = smooth_predictions(merged, sr, hop)
        er = get_er(answers[name], _events(y, hop, sr))
        final_score.append(er)
    if verbose:
        print('FINAL SCORE:', np.mean(final_score))
    return final_score
