"""The evaluation-side chain of metrics.evaluate (metrics.py:30-214).

CPU part: the oracle restatement (oracle/evaluate.py) against hand-worked known answers and
the TF semantics it states.  GPU part: every kernel of k_eval.cu, through the drop-in functions
of challenge_b200.metrics, bit-exact against the oracle on the same inputs, and the whole
``evaluate`` call on synthetic files scored against the reference's ``sample_answer.json``
(tests/golden/sample_answer.json, copied from the reference: it is an INPUT of evaluate)."""
import json
import os
import types

import numpy as np
import pytest

from oracle import evaluate as OE

ANSWERS = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'sample_answer.json')))['task2_answer']


def _track(gt, n_frames, sr=16000, hop=256):
    """0/1 frame track [n_frames, 3] with class c active over [start, end] seconds."""
    y = np.zeros((n_frames, 3), np.float32)
    for c, s, e in gt:
        y[int(s * sr / hop):int((e + 1) * sr / hop), c] = 1
    return y


# ------------------------------------------------------------------------------------------
# CPU: the oracle itself
# ------------------------------------------------------------------------------------------
def test_oracle_frame_windows_pad_end():
    x = np.arange(2 * 7 * 1, dtype=np.float32).reshape(2, 7, 1)
    w = OE.frame_windows(x, frame_length=4, frame_step=3)
    assert w.shape == (3, 2, 4, 1)                        # ceil(7 / 3) windows
    assert np.array_equal(w[0, 0, :, 0], [0, 1, 2, 3])
    assert np.array_equal(w[2, 1, :, 0], [13, 0, 0, 0])   # zero-padded past the end
    # frame_length < step leaves gaps, frame_length > step overlaps
    assert OE.frame_windows(x, 2, 3).shape == (3, 2, 2, 1)


def test_oracle_overlap_average():
    # two windows of 4, hop 2: positions 2, 3 are covered twice
    p = np.array([[[1.], [1.], [3.], [3.]], [[5.], [5.], [7.], [7.]]], np.float32)
    out = OE.overlap_average(p, 2, 6)
    assert np.array_equal(out[:, 0], [1, 1, 4, 4, 7, 7])
    # upsampling repeats each step; a gap is 0 / 0 = NaN as in the reference
    out = OE.overlap_average(np.array([[[2.]], [[4.]]], np.float32), 3, 5, upsample=2)
    assert np.array_equal(out[[0, 1, 3, 4], 0], [2, 2, 4, 4]) and np.isnan(out[2, 0])


def test_oracle_pools_same_padding():
    x = np.zeros((10, 1), np.float32)
    x[4] = 3
    a = OE.avg_pool_same_stride1(x, 3)
    assert np.allclose(a[:, 0], [0, 0, 0, 1, 1, 1, 0, 0, 0, 0])
    x[0] = 3
    a = OE.avg_pool_same_stride1(x, 3)
    assert a[0, 0] == np.float32(1.5)                     # mean over the 2 valid cells only
    m = OE.max_pool_same_stride1(x, 4)                    # SAME: 1 before, 2 after
    assert np.array_equal(m[:, 0], [3, 3, 3, 3, 3, 3, 0, 0, 0, 0])


def test_oracle_events_and_seconds():
    y = np.zeros((100, 3), np.float32)
    y[10:20, 0] = 1
    y[90:, 0] = 1          # still active at the end: closed with len(data)
    y[0:5, 2] = 1          # active from frame 0
    c0, c1, c2 = OE.get_start_end_frame(y)
    assert c0.tolist() == [[10, 19], [90, 99]] and c1.tolist() == [] and c2.tolist() == [[0, 4]]
    rows = OE.output_to_metric(256, 16000)(c0, c1, c2)
    # ((10 + 19) / 2) * 256 / 16000 = 0.232 -> 0 ; ((90 + 99) / 2) * 0.016 = 1.512 -> 1
    assert rows.tolist() == [[0, 0], [0, 1], [2, 0]]
    assert rows.dtype == np.int32


def test_oracle_get_er_hand_worked():
    gt = [[0, 10, 13], [0, 16, 19], [1, 22, 25]]
    # perfect: one prediction inside every ground-truth window
    assert OE.get_er(gt, [[0, 11], [0, 17], [1, 24]])[0] == 0.0
    # nothing predicted: N = 3, answer = 0
    assert OE.get_er(gt, np.zeros((0, 2), np.int32))[0] == 1.0
    # wrong class + one extra: N = 3 + 3, answer = 2 * 1
    er, N, answer = OE.get_er(gt, [[1, 11], [0, 17], [2, 40]])
    assert (N, answer) == (6, 2) and er == 4 / 3
    # a prediction is used once: two ground truths overlapping one prediction
    er, N, answer = OE.get_er([[0, 10, 20], [0, 12, 22]], [[0, 15]])
    assert (N, answer) == (3, 2)
    # greedy in time order: the earlier ground truth takes the earliest prediction
    er, N, answer = OE.get_er([[0, 10, 30], [0, 11, 12]], [[0, 12], [0, 25]])
    assert answer == 2                                    # gt0 takes 12, gt1 is left with 25 (outside)


def test_oracle_chain_scores_perfect_predictions():
    name = sorted(ANSWERS)[0]
    gt = ANSWERS[name]
    T = int((max(e for _, _, e in gt) + 5) * 16000 / 256)
    feats = np.zeros((8, T, 2), np.float32)
    track = _track(gt, T)

    def model(win):
        n_win = win.shape[0]
        p = np.zeros((n_win, 512, 3), np.float32)
        for w in range(n_win):
            seg = track[w * 512:(w + 1) * 512]
            p[w, :len(seg)] = seg
        return p
    er, parts = OE.evaluate_one(feats, model, gt, n_frame=512, n_chan=2)
    # the max pool widens every event by ~2 s; neighbouring same-class events may merge, so
    # the score is not 0, but every predicted event must sit on a ground-truth one
    assert parts['N'] - parts['answer'] <= len(gt)
    assert er <= 1.0


# ------------------------------------------------------------------------------------------
# GPU: k_eval.cu through the drop-in functions
# ------------------------------------------------------------------------------------------
def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope='module')
def M(engine):
    import challenge_b200
    from challenge_b200 import engine as E, metrics
    E._engines[0] = engine
    challenge_b200.set_seed(0)
    return metrics


@pytest.mark.gpu
@pytest.mark.parametrize('T,n_frame,hop', [(1300, 512, 512), (700, 300, 512), (2000, 1024, 512), (511, 512, 512)])
def test_gpu_frame_windows(M, T, n_frame, hop):
    x = np.random.default_rng(T).standard_normal((80, T, 2)).astype(np.float32)
    got = _np(M.frame_windows(x, n_frame, hop))
    assert np.array_equal(got, OE.frame_windows(x, n_frame, hop))


@pytest.mark.gpu
@pytest.mark.parametrize('n_win,n_p,up,hop,L', [(5, 512, 1, 512, 2300), (4, 1024, 1, 512, 2500),
                                              (6, 16, 32, 512, 3000), (3, 300, 1, 512, 1200),
                                              (7, 1536, 1, 512, 4600)])
def test_gpu_overlap_average(M, n_win, n_p, up, hop, L):
    p = np.random.default_rng(n_win).random((n_win, n_p, 3)).astype(np.float32)
    got = _np(M.overlap_average(p, hop, L, up))
    ref = OE.overlap_average(p, hop, L, up)
    assert got.shape == ref.shape
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    two = n_p * up <= 2 * hop          # <= 2 overlapping windows: the sum is order-independent
    if two:
        assert np.array_equal(np.nan_to_num(got), np.nan_to_num(ref))
    else:
        assert np.array_equal(np.nan_to_num(got), np.nan_to_num(ref))   # same ascending order on both sides


@pytest.mark.gpu
@pytest.mark.parametrize('L', [40, 626, 3751])
def test_gpu_smoothing(M, L):
    rng = np.random.default_rng(L)
    # slowly varying tracks that cross 0.5 (plus noise): exercises both pools and the threshold
    t = np.arange(L)[:, None]
    x = (0.5 + 0.5 * np.sin(t / (7.0 + 11 * np.arange(3)[None, :])) + 0.2 * rng.standard_normal((L, 3))).astype(np.float32)
    got = _np(M.smooth_predictions(x))
    ref = OE.smooth(x)
    assert np.array_equal(got, ref)
    assert set(np.unique(got)) <= {0.0, 1.0}


@pytest.mark.gpu
def test_gpu_events_and_seconds(M):
    rng = np.random.default_rng(5)
    for L in (1, 2, 100, 1024, 1025, 5000):
        y = (rng.random((L, 3)) < 0.5).astype(np.float32)
        if L >= 100:
            y = OE.max_pool_same_stride1(OE.avg_pool_same_stride1(y, 5), 3)
            y = (y >= 0.5).astype(np.float32)
        y[-1, 0] = 1                                       # an event that is still open at the end
        cm = M.Challenge_Metric()
        got = [_np(c) for c in cm.get_start_end_frame(y)]
        ref = OE.get_start_end_frame(y)
        for g, r in zip(got, ref):
            assert g.dtype == np.int64 and np.array_equal(g.reshape(-1, 2), r)
        rows_ref = OE.output_to_metric(256, 16000)(*ref)
        rows, n_rows = M._events(y, 256, 16000)
        n = int(_np(n_rows)[0])
        assert n == len(rows_ref)
        assert np.array_equal(_np(rows)[:n][:, [0, 3]], rows_ref)
        assert np.array_equal(_np(M.output_to_metric(256, 16000)(*got)), rows_ref)


@pytest.mark.gpu
def test_gpu_get_er(M):
    rng = np.random.default_rng(9)
    # the hand-worked cases of the oracle test
    gt = [[0, 10, 13], [0, 16, 19], [1, 22, 25]]
    assert M.get_er(gt, [[0, 11], [0, 17], [1, 24]]) == 0.0
    assert M.get_er(gt, np.zeros((0, 2), np.int32)) == 1.0
    assert M.get_er(gt, [[1, 11], [0, 17], [2, 40]]) == 4 / 3
    for trial in range(30):
        m, n = int(rng.integers(1, 40)), int(rng.integers(0, 60))
        s = rng.integers(0, 100, m)
        g = np.stack([rng.integers(0, 3, m), s, s + rng.integers(0, 8, m)], 1)
        p = np.stack([rng.integers(0, 3, n), rng.integers(0, 110, n)], 1).reshape(-1, 2)
        er, N, answer = OE.get_er(g, p)
        got = _np(M.er_counts_events(g, p))
        assert got.tolist() == [N, answer, m], trial
        assert M.get_er(g, p) == er
    with pytest.raises(ZeroDivisionError):
        M.get_er(np.zeros((0, 3), np.int32), [[0, 1]])


@pytest.mark.gpu
@pytest.mark.parametrize('n_frame,v', [(512, 0), (1024, 0), (512, 3)])
def test_gpu_evaluate_matches_oracle(M, n_frame, v):
    """metrics.evaluate end to end on six synthetic 2-channel files named like the reference's
    sample set, scored against sample_answer.json; the model is a stand-in that returns a noisy
    version of the ground-truth track of the windows it is given."""
    import torch
    from challenge_b200 import data_utils as DU, transforms as TR
    rng = np.random.default_rng(11)
    cfg = types.SimpleNamespace(n_chan=2, model_type='vad', n_mels=80, n_frame=n_frame, v=v)
    down = 32 if v in M.label_downsample_model else 1
    wavs, tracks = {}, {}
    for name, gt in ANSWERS.items():
        seconds = max(e for _, _, e in gt) + 3
        wavs[name] = (0.1 * rng.standard_normal((2, seconds * 16000))).astype(np.float32)
    state = {}

    class Model:
        def predict(self, win):
            # win: [n_win, mel, n_frame, n_chan] device tensor
            assert win.shape[1:] == (80, n_frame, 2)
            track = state['track']
            n_win = int(win.shape[0])
            p = np.zeros((n_win, n_frame, 3), np.float32)
            for w in range(n_win):
                seg = track[w * 512:w * 512 + n_frame]
                p[w, :len(seg)] = seg
            p = np.clip(p * 0.8 + 0.1 + 0.05 * state['noise'].standard_normal(p.shape), 0, 1).astype(np.float32)
            if down > 1:
                p = p.reshape(n_win, n_frame // down, down, 3).mean(2).astype(np.float32)
            state['preds'] = p
            state['windows'] = win
            return torch.as_tensor(p)

    scores, ref_scores = [], []
    for name in sorted(wavs):
        gt = ANSWERS[name]
        T = 1 + wavs[name].shape[1] // 256
        state['track'] = _track(gt, T)
        state['noise'] = np.random.default_rng(len(name) + T)
        got = M.evaluate(cfg, Model(), wavs={name: wavs[name]}, answers=ANSWERS)
        scores.append(got[0])
        # the oracle continues from the SAME features / predictions
        feats = DU.log_on_mel(DU.minmax(TR.magphase_to_mel(80)(TR.complex_to_magphase(
            DU.stft_filter(16)(DU.load_wav(wavs[name]))))))
        feats = _np(feats)
        assert feats.shape == (80, T, 2)
        assert np.array_equal(_np(state['windows']), OE.frame_windows(feats, n_frame, 512)[..., :2])
        merged = OE.overlap_average(state['preds'], 512, T, down)
        y = OE.smooth(merged)
        rows = OE.output_to_metric(256, 16000)(*OE.get_start_end_frame(y))
        ref_scores.append(OE.get_er(gt, rows)[0])
    assert scores == ref_scores
    assert all(np.isfinite(s) for s in scores)
    assert min(scores) < 1.0          # the stand-in model does find events
