"""The oracle pinned against every known-answer vector the reference's own tests hold for the
hot path (tests/golden/reference_kats.json), plus independent float64 cross-checks for the
parts the reference never tests (STFT, mel matrix)."""
import json
import os

import numpy as np
import pytest

from oracle import data_utils as D
from oracle import metrics as M
from oracle import pipeline as P
from oracle import transforms as T

KATS = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_kats.json')))


def test_mask_goldens():
    for key in ('mask_axis0', 'mask_axis1'):
        k = KATS[key]
        out = T.mask(np.array(k['org']), axis=k['axis'], max_mask_size=k['max_mask_size'],
                     n_mask=k['n_mask'], draws=k['draws'])
        assert np.array_equal(out, np.array(k['target'])), key


def test_random_shift_golden():
    k = KATS['random_shift']
    out = T.random_shift(np.array(k['org']), axis=k['axis'], width=k['width'], offset=k['offset'])
    assert np.array_equal(out, np.array(k['target']))


def test_log_magphase_golden():
    k = KATS['log_magphase']
    out = T.log_magphase(np.array(k['specs'], np.float64), n_chan=k['n_chan'])
    np.testing.assert_allclose(out, np.array(k['target']), rtol=1e-6, atol=1e-6)


def test_complex_magphase_goldens():
    k = KATS['complex_to_magphase']
    c = np.array(k['complex'], np.float32)
    mp = np.array(k['magphase'], np.float32)
    np.testing.assert_allclose(T.complex_to_magphase(c), mp, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(T.magphase_to_complex(mp), c, rtol=1e-6, atol=1e-6)


def test_minmax_norm_magphase_property():
    rng = np.random.default_rng(0)
    mag = rng.standard_normal((5, 10, 2))
    phase = (2 * rng.random((5, 10, 2)) - 1) * np.pi
    out = T.minmax_norm_magphase(np.concatenate([mag, phase], -1))
    np.testing.assert_allclose(out.min(axis=(1, 2)), 0, atol=1e-6)
    np.testing.assert_allclose(out.max(axis=(1, 2)), 1, atol=1e-6)


def test_er_score_golden():
    k = KATS['er_score']
    g = np.zeros([k['batch'], k['frames'], k['classes']])
    p = np.zeros_like(g)
    for c, s, e in k['gt']:
        g[:, s:e, c] = 1
    for c, t in k['predict']:
        p[:, t - 2:t + 2, c] = 1
    er = M.er_score(smoothing=False)(g, p)
    assert np.float32(er.mean()) == np.float32(k['mean_er'])
    nt, npd, co = M.er_parts(g, p)
    assert nt.tolist() == [5, 5] and npd.tolist() == [5, 5] and co.tolist() == [2, 2]


def test_pipeline_shapes():
    s = KATS['shapes']['merge_complex_specs']
    rng = np.random.default_rng(1)
    F, C, K, T_ = s['freq'], s['chan'], s['n_classes'], s['n_frame']
    bg = rng.standard_normal((F, s['bg_frames'], C)).astype(np.float32)
    voices = rng.standard_normal((s['n_voices'], F, T_, C)).astype(np.float32)
    lens = rng.integers(1, T_, size=s['n_voices'])
    voices *= (np.arange(T_)[None, :] < lens[:, None])[:, None, :, None]
    labels = np.eye(K, dtype=np.float32)[rng.integers(1, T_, size=s['n_voices'])]
    noises = rng.standard_normal((s['n_noises'], F, T_, C)).astype(np.float32)
    # min_ratio 2/3: pad = 10 - int(6.67) = 4 -> len 18 -> offsets in [0, 8)
    draws = dict(bg_offset=3, n_voices=3, voice_u=[0.5, 1.0, 1.5], voice_offset=[0, 7, 3],
                 n_noises=1, noise_u=[0.3], noise_offset=[5])
    spec, lab = P.merge_complex_specs(bg, (voices, labels), noises, n_frame=T_, n_classes=K,
                                      draws=draws)
    assert spec.shape == (F, T_, C) and lab.shape == (s['n_voices'], T_, K)
    # background tiling: frame t of the crop is bg[(t + 3) % 8] wherever nothing was added
    assert set(np.unique(lab)) <= {0.0, 1.0}
    mel = T.magphase_to_mel(80)(np.zeros(KATS['shapes']['magphase_to_mel']['in'], np.float32))
    assert list(mel.shape) == KATS['shapes']['magphase_to_mel']['out']
    assert T.magphase_to_mel(80)(np.zeros((257, 100, 4), np.float32)).shape == (80, 100, 2)


def test_make_pipeline_shapes():
    s = KATS['shapes']['make_pipeline']
    rng = np.random.default_rng(2)
    nf, K = s['n_frame'], 30
    bgs = [rng.standard_normal((257, rng.integers(1, nf * 2), 4)).astype(np.float32)
           for _ in range(s['n_bg'])]
    voices = [rng.standard_normal((257, rng.integers(1, nf // 2), 4)).astype(np.float32)
              for _ in range(s['n_voice'])]
    labels = np.eye(K, dtype=np.float32)[rng.integers(K, size=s['n_voice'])]
    noises = [rng.standard_normal((257, rng.integers(1, nf // 2), 4)).astype(np.float32)
              for _ in range(s['n_noise'])]
    V, Mx = s['max_voices'], s['max_noises']
    for e in range(3):
        vs = rng.permutation(s['n_voice'])
        ns = rng.permutation(s['n_noise'])
        bs = rng.permutation(s['n_bg'])
        vP = max(voices[i].shape[1] for i in vs[e * V:(e + 1) * V])
        nP = max(noises[i].shape[1] for i in ns[e * Mx:(e + 1) * Mx])
        vlen = vP + 2 * (nf - int(np.int32(np.float32(2 / 3) * np.float32(vP))))
        nlen = nP + 2 * (nf - int(np.int32(np.float32(0.5) * np.float32(nP))))
        bgT = bgs[bs[e]].shape[1]
        tiled = bgT * ((nf + bgT - 1) // bgT)
        draws = dict(bg_offset=int(rng.integers(0, tiled - nf + 1)), n_voices=2,
                     voice_u=[0.1, 0.2], voice_offset=[0, vlen - nf - 1], n_noises=1,
                     noise_u=[1.0], noise_offset=[nlen - nf])
        spec, lab = P.make_pipeline_element(bgs, voices, labels, noises, e, bg_stream=bs,
                                            voice_stream=vs, noise_stream=ns, n_frame=nf,
                                            max_voices=V, max_noises=Mx, n_classes=K, draws=draws)
        assert spec.shape == (257, nf, 4) and lab.shape == (V, nf, K)


def test_phase_vocoder_identity_and_shapes():
    k = KATS['shapes']['phase_vocoder']
    x = np.random.default_rng(3).standard_normal(k['in']).astype(np.float32)
    assert np.array_equal(T.phase_vocoder(x, 1.), x)
    for rate in k['rates']:
        assert list(T.phase_vocoder(x, rate).shape) == [257, int(np.ceil(100 / rate)), 6]


def test_empty_offset_range_raises():
    rng = np.random.default_rng(4)
    bg = rng.standard_normal((257, 12, 2)).astype(np.float32)
    voices = rng.standard_normal((2, 257, 10, 2)).astype(np.float32)
    labels = np.eye(3, dtype=np.float32)[[0, 1]]
    draws = dict(bg_offset=0, n_voices=1, voice_u=[0.1], voice_offset=[0])
    try:
        P.merge_complex_specs(bg, (voices, labels), None, n_frame=10, min_ratio=1, draws=draws)
    except P.InvalidArgumentError:
        return
    raise AssertionError('pipeline.py:68-69 must raise when len == n_frame')


def test_stft_against_float64_restatement():
    """PARITY UNPINNED in the reference (no test touches load_wav); the oracle's STFT is the
    reference's own torchaudio call, cross-checked here against SURVEY A.1 in float64."""
    rng = np.random.default_rng(5)
    for C, N in [(2, 16000), (4, 5000), (1, 300)]:
        x = D.normalize((rng.standard_normal((C, N)) * 0.1).astype(np.float32))
        s = D.stft(x)
        s64 = D.stft_f64(x)
        assert s.shape == (C, 257, 1 + N // 256)
        assert np.abs(s - s64).max() / np.abs(s64).max() < 2e-6
        assert not s.imag[:, 0].any() and not s.imag[:, 256].any()
    lay = D.spec_layout(s)
    assert lay.shape == (257, 1 + N // 256, 2 * C)
    assert np.array_equal(lay[..., :C], s.real.transpose(1, 2, 0))
    assert np.array_equal(lay[..., C:], s.imag.transpose(1, 2, 0))


def test_mel_matrix_structure():
    """PARITY UNPINNED by the reference (transforms_test.py:45-55 is shape-only).  The TF-2.2
    semantics restated in fp32 (oracle) are pinned to tests/golden/mel_matrix_80.json, built by
    scripts/make_mel_golden.py from the published formula in float64, independently of oracle/
    and of the product: support (231 non-zeros, rows 5..121, taps per filter), every column and
    row sum, every entry of 12 columns."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'mel_matrix_80.json')))
    w = T.linear_to_mel_weight_matrix(80, 257, 16000)
    assert w.shape == tuple(g['shape']) and w.dtype == np.float32
    assert (w != 0).sum() == g['nnz'] == 231 and not w[0].any()
    rows = np.nonzero(w.sum(1))[0]
    assert rows[0] == g['first_row'] == 5 and rows[-1] == g['last_row'] == 121
    assert (w != 0).sum(1).max() == g['max_taps_per_row'] == 2
    assert (w != 0).sum(0).tolist() == g['taps_per_column']
    # fp32 build (mel values ~2000 carry 1.2e-4 of rounding, a triangle is ~24 mel wide) vs float64
    assert np.abs(w.sum(0) - np.array(g['column_sums'])).max() < 4e-5
    assert np.abs(w.sum(1) - np.array(g['row_sums'])).max() < 1e-5
    assert len(g['entries']) >= 30
    for f, j, v in g['entries']:
        assert abs(float(w[f, j]) - v) < 2e-5, (f, j)   # measured 1.06e-5 at (40, 40)
    # and a second float64 rebuild inline
    lin = np.linspace(0, 8000, 257)[1:]
    h2m = lambda f: 1127.0 * np.log1p(np.asarray(f, np.float64) / 700.0)
    e = np.linspace(h2m(125.0), h2m(3800.0), 82)
    sm = h2m(lin)[:, None]
    w64 = np.maximum(0, np.minimum((sm - e[None, :-2]) / (e[None, 1:-1] - e[None, :-2]),
                                   (e[None, 2:] - sm) / (e[None, 2:] - e[None, 1:-1])))
    assert np.abs(w[1:] - w64).max() < 2e-5


def test_label_downsample_semantics():
    y = np.zeros((2, 70, 3), np.float32)
    y[0, 0:20, 0] = 1        # window 0 (valid cells 0..18 of [-13, 19)) fully set
    y[1, 40:56, 1] = 1
    _, out = D.label_downsample(32)(None, y)
    assert out.shape == (2, 3, 3)       # ceil(70/32) = 3 windows; [:32] keeps both samples
    # pad_left = (96 - 70) // 2 = 13: windows [-13,19) [19,51) [51,83)
    assert out[0, 0, 0] == 1 and out[0, 1, 0] == 0
    assert out[1, 1, 1] == 0 and out[1, 2, 1] == 0      # 11/32 and 5/19 < 0.5
    _, out = D.label_downsample(2)(None, np.ones((5, 4, 3), np.float32))
    assert out.shape == (2, 2, 3)       # the [:resolution] slice hits the BATCH axis (quirk)


def test_f1_and_cos_sim():
    y = np.zeros((2, 8, 3), np.float32)
    y[0, 2:5, 0] = 1
    p = np.zeros_like(y)
    p[0, 3:7, 0] = 0.9
    p[1, 0, 2] = 0.5            # not > 0.5: no false positive
    assert M.f1_counts(y, p) == (2, 2, 1)
    f = M.f1_score()
    assert np.isclose(f(y, p), 2 * 0.5 * (2 / 3) / (0.5 + 2 / 3))
    assert M.f1_counts(y, p) == (2, 2, 1) and (f.tp, f.fp, f.fn) == (2, 2, 1)
    f(y, p)
    assert (f.tp, f.fp, f.fn) == (4, 4, 2)      # never reset (metrics.py:291-297)
    cs = M.cos_sim(y, y)
    assert np.isclose(cs[0], -1.0) and cs[1] == 0


@pytest.mark.parametrize('orig,n', [(44100, 22051), (48000, 30001), (8000, 9000), (22050, 12345)])
def test_oracle_resampler_is_pinned_to_torchaudio(orig, n):
    """The reference calls torchaudio.compliance.kaldi.resample_waveform (data_utils.py:20-21), gone
    from torchaudio 2.x; the same filter (Hann-windowed sinc, width 6, cutoff 0.99 Nyquist, zero
    padding, ceil(len * new / orig) outputs) is torchaudio.functional.resample, run here as the
    pin of the oracle's restatement of the Kaldi port."""
    torch = pytest.importorskip('torch')
    ta = pytest.importorskip('torchaudio')
    from oracle import data_utils as OD
    rng = np.random.default_rng(orig)
    wav = rng.standard_normal((2, n)).astype(np.float32)
    got = OD.resample_waveform(wav, orig, 16000)
    ref = ta.functional.resample(torch.from_numpy(wav), orig, 16000).numpy()
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max()


def _synth_index_level(bg, voices, vlabels, noises, d, b, K=3):
    """Independent restatement of sample synthesis from the INDEX-LEVEL semantics (SURVEY.md appendix
    A.2), frame by frame in float64, sharing no code with oracle/pipeline.py (which pads, slices and
    adds whole tensors like the reference does): out[f,t] = BG[f,(t+o_b) mod bgT] + sum over kept
    voices of g_v X_v[f, t+off_v-s_v] + sum over noises; labels and keep flags by the same indices."""
    T = d.n_frame
    V, M = d.max_voices, d.max_noises
    BG = bg[int(d.bg_id[b])].astype(np.float64)
    bgT = BG.shape[1]
    out = np.empty((BG.shape[0], T, BG.shape[2]), np.float64)
    for t in range(T):
        out[:, t] = BG[:, (t + int(d.bg_offset[b])) % bgT]
    label = np.zeros((V, T, K), np.float64)
    keep = [0] * V
    if V:
        ids = [int(i) for i in d.voice_id[b]]
        vP = max(voices[i].shape[1] for i in ids)                      # padded_batch: the group's longest
        pad = T - int(np.int32(np.float32(d.min_ratio) * np.float32(vP)))
        s = pad if pad > 0 else 0
        nv = int(d.n_voices[b]) if V > 1 else 1
        L = np.zeros((T, K), np.float64)
        for v in range(nv):
            X = voices[ids[v]].astype(np.float64)
            kT = X.shape[1]
            act = (voices[ids[v]].max(axis=(0, 2)) > 0)
            cand = np.zeros((T, K), np.float64)
            src = [t + int(d.voice_offset[b][v]) - s for t in range(T)]
            for t, k in enumerate(src):
                if 0 <= k < kT and act[k]:
                    cand[t] = np.asarray(vlabels[ids[v]], np.float64)
            ok = not ((L + cand) >= 2).any()
            keep[v] = int(ok)
            if ok:
                g = float(d.voice_gain[b][v])
                for t, k in enumerate(src):
                    if 0 <= k < kT:
                        out[:, t] += g * X[:, k]
                L += cand
                label[v] = cand
    if M:
        ids = [int(i) for i in d.noise_id[b]]
        nP = max(noises[i].shape[1] for i in ids)
        pad = T - int(np.int32(np.float32(d.min_noise_ratio) * np.float32(nP)))
        s = pad if pad > 0 else 0
        for n in range(int(d.n_noises[b])):
            X = noises[ids[n]].astype(np.float64)
            g = float(d.noise_gain[b][n])
            for t in range(T):
                k = t + int(d.noise_offset[b][n]) - s
                if 0 <= k < X.shape[1]:
                    out[:, t] += g * X[:, k]
    return out, label, keep


def test_merge_complex_specs_against_an_index_level_restatement():
    """PARITY UNPINNED by the reference (pipeline_test.py checks shapes only): the oracle's tensor-level
    restatement of merge_complex_specs (pad / slice / add like pipeline.py:41-106) is cross-checked
    against an independent frame-by-frame float64 implementation of the index semantics -- tiled
    background crop, padded-group length, pad / offset arithmetic, activity, same-class overlap
    rejection, gains -- on random draws: values to 1e-6, labels and keep flags exactly."""
    from challenge_b200.plan import draw_batch
    from challenge_b200.synth import synthetic_banks
    from oracle import chain
    for seed, (C, T, V, M, mr) in enumerate([(2, 90, 4, 2, 2 / 3), (1, 60, 3, 0, 1.0), (2, 150, 5, 3, 1.0), (3, 40, 2, 2, 2 / 3)]):
        bgs, voices, labels, noises = synthetic_banks(300 + seed, C, n_bg=3, n_voice=8, n_noise=3, bg_seconds=0.9,
                                                      lo_s=0.2, hi_s=1.2)
        ob, ov, on = chain.OracleBank(bgs), chain.OracleBank(voices), chain.OracleBank(noises)
        bf = np.array([s.shape[1] for s in ob.specs], np.int32)
        vf = np.array([s.shape[1] for s in ov.specs], np.int32)
        nf = np.array([s.shape[1] for s in on.specs], np.int32)
        try:
            d = draw_batch(np.random.default_rng(seed), 6, T, bf, vf, nf if M else None, max_voices=V, max_noises=M,
                           snr=-20, min_ratio=mr)
        except ValueError:
            continue
        n_checked = 0
        for b in range(d.batch):
            dbg = {}
            spec, lab = chain.synth_clip(ob, ov, labels, on if M else None, d, b, debug=dbg)
            ref, ref_lab, ref_keep = _synth_index_level(ob.specs, ov.specs, labels, on.specs if M else None, d, b)
            assert np.abs(spec - ref).max() <= 1e-6 * np.abs(ref).max(), (seed, b)
            assert np.array_equal(lab, ref_lab.astype(np.float32)), (seed, b)
            assert list(dbg.get('no_overlap', [0] * V))[:int(d.n_voices[b]) if V > 1 else 1] == \
                ref_keep[:int(d.n_voices[b]) if V > 1 else 1], (seed, b)
            n_checked += 1
        assert n_checked == d.batch
