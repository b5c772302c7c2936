"""GPU parity tests proper: the CUDA path through the C ABI vs the CPU oracle on the same
seeded inputs and the same host draws.  Bars (BASELINE.json north_star): bit-exact masks /
labels / keep flags / metric counts; float features within normalised max error 1e-4
(max|a-b| / max|b|); phase as circular error gated on magnitude."""
import numpy as np
import pytest

from conftest import logmag_report, nmax_err, phase_err, phase_report, rel_err_gated

pytestmark = pytest.mark.gpu

TOL = 1e-4
# Phase and log-magnitude bars.  Both sides run the FFT in fp32, so re / im carry an absolute error
# of ~3e-7 of max|X| and a cell of magnitude m has a phase (and log-magnitude) error of ~3e-7 max / m:
# measured on B200 (scripts/parity_report.py, cfg3 shape, three draws) 4.1e-5 .. 8.4e-5 rad and
# 5.8e-5 .. 8.9e-5 in the log at the 1e-3 gate, 1.0e-5 at a 1e-2 gate, 2.7e-7 magnitude-weighted.
PHASE_TOL = 2e-4          # rad, cells with |X| > 1e-3 max|X|
PHASE_WEIGHTED_TOL = 2e-6  # |dphi| |X| / max|X|, every cell
LOGMAG_TOL = 2e-4         # |d log|, cells with |X| > 1e-3 max|X|
NOISE = 2e-6              # of max|X|: an imaginary part below this has no resolvable sign


def _check_phase(mag_ref, ph, ph_ref, c_got=None, c_ref=None):
    """circular phase error by magnitude gate + the count of +-pi sign mismatches.
    A mismatch is only legitimate where the imaginary part is rounding noise on both sides (e.g.
    the reflect-padded edge frames, whose spectrum is real up to rounding); cells of exactly zero
    magnitude (masked) must agree up to the signs of noise-level components."""
    r = phase_report(mag_ref, ph, ph_ref)
    assert r['gate1e-3'] < PHASE_TOL, r
    assert r['weighted'] < PHASE_WEIGHTED_TOL, r
    C = mag_ref.shape[-1]
    if c_got is not None:
        raw = ph.astype(np.float64) - ph_ref
        flips = np.abs(raw) > np.pi
        mx = np.abs(c_ref).max()
        im_g, im_r = np.abs(c_got[..., C:][flips]), np.abs(c_ref[..., C:][flips])
        assert (im_g <= NOISE * mx).all() and (im_r <= NOISE * mx).all(), (r, im_g.max(), im_r.max())
    if r['zero_cells']:
        assert r['zero_cells_differ'] <= 2e-3 * r['zero_cells'] + 8, r
    return r



def _draw(w, B, T, C_remap=0, seed=0, masks=True, V=7, M=2, min_ratio=1, merge_extra=0):
    from challenge_b200.plan import draw_batch
    rng = np.random.default_rng(seed)
    return draw_batch(rng, B, T, w.bg_frames, w.voice_frames, w.noise_frames, max_voices=V,
                      max_noises=M, snr=-20, min_ratio=min_ratio,
                      n_time_masks=6 if masks else 0, n_freq_masks=1 if masks else 0,
                      merge_extra=merge_extra)


def _oracle(w, d, **kw):
    from oracle import chain
    return chain.dataset_batch(w.o_bg, w.o_voice, w.labels, w.o_noise, d, **kw)


def test_stft_matches_reference_call(engine):
    """load_wav (data_utils.py:9-29): normalize + Spectrogram(512) + [F,T,2C] layout."""
    from oracle import data_utils as D
    rng = np.random.default_rng(3)
    for C, N in [(2, 160000), (4, 40000), (1, 5000), (3, 12345), (2, 257)]:
        wav = (rng.standard_normal((C, N)) * 0.1).astype(np.float32)
        got = engine.stft(wav).cpu().numpy()
        ref = D.load_wav_array(wav)
        assert got.shape == ref.shape == (257, 1 + N // 256, 2 * C)
        assert nmax_err(got, ref) < TOL
        # DC and Nyquist imaginary parts are exactly +0.0 like torch.stft
        assert not np.any(got[0, :, C:]) and not np.any(np.signbit(got[0, :, C:]))
        assert not np.any(got[256, :, C:]) and not np.any(np.signbit(got[256, :, C:]))
    got = engine.stft(wav, normalize=False).cpu().numpy()
    assert nmax_err(got, D.load_wav_array(wav, do_normalize=False)) < TOL


def test_voice_activity_bit_exact(engine, workload_factory):
    """pipeline.py:55 -- frame active iff any STFT coefficient > 0."""
    w = workload_factory(2)
    for i in range(len(w.voices)):
        assert np.array_equal(engine.voice_activity(i), w.o_voice.activity(i)), i
    assert any(0 in engine.voice_activity(i) for i in range(len(w.voices)))


def test_cfg1_logmel_no_augmentation(engine, workload_factory):
    """BASELINE config 1: 2-ch 10 s clips -> log-mel, no voices / noises / masks."""
    from challenge_b200 import _lib as L
    from challenge_b200.plan import draw_batch
    w = workload_factory(2)
    d = draw_batch(np.random.default_rng(1), 4, 626, w.bg_frames)
    engine.upload_plan(d)
    for mode, name in [(L.FEAT_MEL, 'mel'), (L.FEAT_LOGMEL, 'logmel'),
                       (L.FEAT_LOGMEL_MINMAX, 'logmel_minmax')]:
        got = engine.features(mode).cpu().numpy()
        ref, _, _, _ = _oracle(w, d, mode=name)
        assert got.shape == ref.shape == (4, 80, 626, 2)
        assert nmax_err(got, ref) < TOL, name


@pytest.mark.parametrize('T', [626, 512, 100])
def test_cfg2_mix_masks_labels(engine, workload_factory, T):
    """BASELINE config 2: noise mixing at sampled SNR + time/freq masks + labels."""
    from challenge_b200 import _lib as L
    w = workload_factory(2)
    d = _draw(w, 8, T, seed=T)
    engine.upload_plan(d)
    frame, vtk, keep = engine.labels(want_vtk=True)
    ref_c, ref_y, ref_vtk, ref_keep = _oracle(w, d, mode='complex')
    assert np.array_equal(keep.cpu().numpy(), np.stack(ref_keep))
    assert np.array_equal(frame.cpu().numpy(), ref_y)
    assert np.array_equal(vtk.cpu().numpy(), np.stack(ref_vtk))
    got_c = engine.features(L.FEAT_COMPLEX).cpu().numpy()
    assert got_c.shape == ref_c.shape
    assert nmax_err(got_c, ref_c) < TOL
    # masked cells (transforms.py:12-40) are exact zeros in both, and nothing else is zeroed:
    # outside the mask and the DC / Nyquist imaginary parts an exact 0 may only stand where
    # the reference holds rounding noise (frame 0 of an un-shifted source is symmetric about
    # its centre after reflect padding, so its imaginary part is analytically 0)
    m = np.zeros(got_c.shape, bool)
    for b in range(d.batch):
        for size, off in d.time_masks[b]:
            m[b, :, off:off + size] = True
        for size, off in d.freq_masks[b]:
            m[b, off:off + size] = True
    assert not got_c[m].any() and not ref_c[m].any()
    m[:, 0, :, 2:] = m[:, 256, :, 2:] = True
    assert not ref_c[:, 0, :, 2:].any() and not got_c[:, 0, :, 2:].any()
    assert not ref_c[:, 256, :, 2:].any() and not got_c[:, 256, :, 2:].any()
    stray = (got_c == 0) & ~m
    assert np.abs(ref_c[stray]).max(initial=0) < 1e-6 * np.abs(ref_c).max()
    stray = (ref_c == 0) & ~m
    assert np.abs(got_c[stray]).max(initial=0) < 1e-6 * np.abs(ref_c).max()
    for mode, name in [(L.FEAT_MEL, 'mel'), (L.FEAT_LOGMEL_MINMAX, 'logmel_minmax')]:
        got = engine.features(mode).cpu().numpy()
        ref = _oracle(w, d, mode=name)[0]
        assert nmax_err(got, ref) < TOL, name
        if name == 'mel':      # linear features: the element-wise relative error is well posed above a gate
            print('cfg2 mel: gated element-wise relative error %.2e' % rel_err_gated(got, ref))
            assert rel_err_gated(got, ref) < 1e-4      # measured 1.3e-5 .. 1.6e-5: north_star's bar holds element-wise
        else:                  # min-max log-mel lives in [log 1e-8, 0]: absolute error per clip
            assert np.abs(got.astype(np.float64) - ref).max() < 2e-4


def test_cfg3_four_channel_magphase_labels(engine, workload_factory):
    """BASELINE config 3: 4-ch clips, magnitude + phase features and frame labels."""
    from challenge_b200 import _lib as L
    w = workload_factory(4, seed=20203, n_bg=2, n_voice=16, n_noise=4)
    d = _draw(w, 4, 626, seed=33)
    engine.upload_plan(d)
    frame, _, keep = engine.labels()
    ref, ref_y, _, ref_keep = _oracle(w, d, mode='magphase')
    assert np.array_equal(frame.cpu().numpy(), ref_y)
    assert np.array_equal(keep.cpu().numpy(), np.stack(ref_keep))
    got = engine.features(L.FEAT_MAGPHASE).cpu().numpy()
    assert got.shape == ref.shape == (4, 257, 626, 8)
    assert nmax_err(got[..., :4], ref[..., :4]) < TOL
    assert rel_err_gated(got[..., :4], ref[..., :4]) < 5e-4       # element-wise, |X| > 1e-3 max|X|
    got_c = engine.features(L.FEAT_COMPLEX).cpu().numpy()     # both channel pairs stored as 32-byte cells
    ref_c = _oracle(w, d, mode='complex')[0]
    assert got_c.shape == ref_c.shape == (4, 257, 626, 8)
    assert nmax_err(got_c, ref_c) < TOL
    rep = _check_phase(ref[..., :4], got[..., 4:], ref[..., 4:], got_c, ref_c)
    print('cfg3 phase report:', rep)
    for ch in range(8):                                       # every channel lands in its own column
        assert nmax_err(got_c[..., ch], ref_c[..., ch]) < 10 * TOL, ch
    got = engine.features(L.FEAT_LOG_MAGPHASE).cpu().numpy()
    ref = _oracle(w, d, mode='log_magphase')[0]
    lr = logmag_report(got[..., :4], ref[..., :4])
    assert lr['linear_nmax'] < TOL and lr['gate1e-3'] < LOGMAG_TOL, lr
    assert phase_report(np.exp(ref[..., :4]), got[..., 4:], ref[..., 4:])['gate1e-3'] < PHASE_TOL   # phase passes through
    got = engine.features(L.FEAT_LOGMEL_MINMAX).cpu().numpy()
    assert nmax_err(got, _oracle(w, d, mode='logmel_minmax')[0]) < TOL


def test_wide_mel_matrix_uses_all_bins(engine, workload_factory):
    """A mel matrix whose support reaches above bin 128 (80 bins, 125-7600 Hz) runs the
    all-bins variant of the fused mel epilogue; the default matrix is restored afterwards."""
    from challenge_b200 import _lib as L
    from oracle import transforms as OT
    w = workload_factory(2)
    d = _draw(w, 4, 300, seed=77)
    try:
        engine.set_mel(80, upper_edge_hertz=7600.0)
        engine.upload_plan(d)
        mel_fn = OT.magphase_to_mel(80, upper_edge_hertz=7600.0)
        for mode, name in [(L.FEAT_MEL, 'mel'), (L.FEAT_LOGMEL_MINMAX, 'logmel_minmax')]:
            got = engine.features(mode).cpu().numpy()
            ref = _oracle(w, d, mode=name, mel_fn=mel_fn)[0]
            assert got.shape == ref.shape == (4, 80, 300, 2)
            assert nmax_err(got, ref) < TOL, name
        # a matrix with filters wider than the fused epilogue takes is refused, not mis-computed
        engine.set_mel(40, lower_edge_hertz=80.0, upper_edge_hertz=7600.0)
        engine.upload_plan(d)
        with pytest.raises(NotImplementedError):
            engine.features(L.FEAT_MEL)
    finally:
        engine.set_mel(80)


def test_background_tiling_and_short_banks(engine, workload_factory):
    """Backgrounds shorter than n_frame are tiled then cropped (pipeline.py:29-35)."""
    from challenge_b200 import _lib as L
    w = workload_factory(2, seed=7, n_bg=3, n_voice=12, n_noise=4, bg_seconds=1.7)
    d = _draw(w, 6, 300, seed=5, min_ratio=2 / 3)
    engine.upload_plan(d)
    frame, _, keep = engine.labels()
    ref, ref_y, _, ref_keep = _oracle(w, d, mode='complex')
    assert np.array_equal(frame.cpu().numpy(), ref_y)
    assert np.array_equal(keep.cpu().numpy(), np.stack(ref_keep))
    assert nmax_err(engine.features(L.FEAT_COMPLEX).cpu().numpy(), ref) < TOL


def test_channel_remaps_and_filter(engine, workload_factory):
    """stereo_mono / random_merge_aug / stft_filter (data_utils.py:79-136) fused in the epilogue."""
    from challenge_b200 import _lib as L
    w = workload_factory(2)
    d = _draw(w, 3, 200, seed=11, merge_extra=2)
    engine.upload_plan(d, stft_filter=3, chan_remap=L.REMAP_STEREO_MONO)
    got = engine.features(L.FEAT_COMPLEX).cpu().numpy()
    ref = _oracle(w, d, mode='complex', remap='stereo_mono', stft_filter=3)[0]
    assert got.shape == ref.shape == (3, 257, 200, 6)
    assert nmax_err(got, ref) < TOL
    assert not np.any(got[:, 1:4])
    engine.upload_plan(d, chan_remap=L.REMAP_MERGE_AUG, n_out_chan=4)
    got = engine.features(L.FEAT_MAGPHASE).cpu().numpy()
    ref = _oracle(w, d, mode='magphase', remap='merge_aug', n_out_chan=4)[0]
    assert got.shape == ref.shape == (3, 257, 200, 8)
    assert nmax_err(got[..., :4], ref[..., :4]) < TOL
    _check_phase(ref[..., :4], got[..., 4:], ref[..., 4:])


def test_empty_offset_range_raises_like_reference(engine, workload_factory):
    """pipeline.py:68-69: int-uniform with maxval == 0 raises when the padded voice length
    equals n_frame with min_ratio=1."""
    from challenge_b200.errors import InvalidArgumentError
    w = workload_factory(2)
    d = _draw(w, 2, 626, seed=2)
    d.n_frame = int(w.voice_frames[d.voice_id[0]].max())
    d.bg_offset[:] = 0
    with pytest.raises(InvalidArgumentError):
        engine.upload_plan(d)


def _blocky(rng, B, T, K, p=0.4):
    y = np.zeros((B, T, K), np.float32)
    for b in range(B):
        for c in range(K):
            t = 0
            while t < T:
                n = int(rng.integers(1, 60))
                if rng.random() < p:
                    y[b, t:t + n, c] = 1
                t += n
    return y


def test_metric_counts_bit_exact(engine):
    """er_score / f1_score counting (metrics.py:217-298) incl. the reference's golden ER."""
    from oracle import metrics as M
    # golden: metrics_test.py:12-25  => mean ER == 1.2
    gt = [[0, 0, 10], [2, 0, 20], [1, 15, 30], [2, 31, 40], [1, 32, 35]]
    pr = [[1, 5], [1, 19], [2, 32], [2, 38], [0, 38]]
    g = np.zeros([2, 40, 3], np.float32)
    p = np.zeros([2, 40, 3], np.float32)
    for c, s, e in gt:
        g[:, s:e, c] = 1
    for c, t in pr:
        p[:, t - 2:t + 2, c] = 1
    triples, _, er = engine.metric_counts(g, p)
    assert triples.cpu().numpy().tolist() == [[5, 5, 2], [5, 5, 2]]
    assert float(er.cpu().numpy().mean()) == np.float32(1.2)
    rng = np.random.default_rng(0)
    for (B, T) in [(64, 626), (7, 33), (3, 1), (16, 512), (5, 31), (5, 32)]:
        yt = _blocky(rng, B, T, 3)
        yp = np.clip(yt + rng.normal(0, 0.35, yt.shape), 0, 1).astype(np.float32)
        triples, tpfpfn, er = engine.metric_counts(yt, yp)
        nt, npd, co = M.er_parts(yt, yp)
        assert np.array_equal(triples.cpu().numpy(), np.stack([nt, npd, co], 1))
        assert tuple(tpfpfn.cpu().numpy().tolist()) == M.f1_counts(yt, yp)
        ref_er = M.er_from_parts(nt, npd, co)
        assert np.array_equal(er.cpu().numpy(), ref_er, equal_nan=True)
    # F1 state accumulates across calls (metrics.py:291-297 closure)
    triples, tpfpfn, _ = engine.metric_counts(yt, yp, tpfpfn=tpfpfn)
    assert tuple(tpfpfn.cpu().numpy().tolist()) == tuple(2 * v for v in M.f1_counts(yt, yp))


@pytest.mark.parametrize('stft_filter', [0, 3])
def test_fixed_epilogue_instances_equal_the_generic_kernel(engine, workload_factory, monkeypatch, stft_filter):
    """The 2-channel mel modes run kernel instances whose epilogue switches are template constants
    (k_fused<FM_MEL, 4, EPI_*>); IRIS_NO_FIXED_EPI routes the same launch through the generic
    instance.  Same arithmetic in the same order: the outputs must be bit-identical, with masks,
    stft_filter and a clip count that makes every CTA change clips mid-claim."""
    from challenge_b200 import _lib as L
    w = workload_factory(2)
    d = _draw(w, 24, 626, seed=909)
    engine.upload_plan(d, stft_filter=stft_filter)
    engine.labels()
    for mode in (L.FEAT_MEL, L.FEAT_LOGMEL, L.FEAT_LOGMEL_MINMAX):
        monkeypatch.delenv('IRIS_NO_FIXED_EPI', raising=False)
        fixed = engine.features(mode).cpu().numpy()
        monkeypatch.setenv('IRIS_NO_FIXED_EPI', '1')
        generic = engine.features(mode).cpu().numpy()
        monkeypatch.delenv('IRIS_NO_FIXED_EPI', raising=False)
        assert np.array_equal(fixed, generic), mode
    ref = _oracle(w, d, mode='logmel_minmax', stft_filter=stft_filter)[0]
    assert nmax_err(engine.features(L.FEAT_LOGMEL_MINMAX).cpu().numpy(), ref) < TOL


def test_full_size_properties_cfg2(engine, workload_factory):
    """BASELINE config 2 at full batch (256): size-independent properties + spot parity."""
    from challenge_b200 import _lib as L
    w = workload_factory(2)
    B = 256
    d = _draw(w, B, 626, seed=2024)
    engine.upload_plan(d)
    frame, vtk, keep = engine.labels(want_vtk=True)
    x = engine.features(L.FEAT_LOGMEL_MINMAX)
    x2 = engine.features(L.FEAT_LOGMEL_MINMAX)
    assert bool((x == x2).all()), 'not deterministic'
    xn = x.cpu().numpy()
    assert xn.shape == (B, 80, 626, 2) and np.isfinite(xn).all()
    # min-max then log: every clip spans exactly [log(1e-8), log(1 + 1e-8)]
    assert np.allclose(xn.reshape(B, -1).min(1), np.log(np.float32(1e-8)), rtol=0, atol=1e-5)
    assert np.allclose(xn.reshape(B, -1).max(1), 0, atol=1e-6)
    # time-masked frames are the clip minimum everywhere
    for b in range(0, B, 37):
        for size, off in d.time_masks[b]:
            assert np.allclose(xn[b, :, off:off + size], np.log(np.float32(1e-8)), rtol=0, atol=1e-5)
    fr = frame.cpu().numpy()
    assert set(np.unique(fr)) <= {0.0, 1.0}
    assert np.array_equal(vtk.cpu().numpy().sum(1), fr)
    kp = keep.cpu().numpy()
    assert np.all(kp[:, 0] == 1)          # the first voice can never collide
    assert not kp[np.arange(7)[None, :] >= d.n_voices[:, None]].any()
    clips = sorted(set([0, 255] + np.random.default_rng(7).choice(B, 32, replace=False).tolist()))
    ref, ref_y, _, ref_keep = _oracle(w, d, mode='logmel_minmax', clips=clips)
    assert nmax_err(xn[clips], ref) < TOL
    for i, b in enumerate(clips):                     # every clip on its own scale
        assert nmax_err(xn[b], ref[i]) < TOL, b
    assert np.array_equal(fr[clips], ref_y)
    assert np.array_equal(kp[clips], np.stack(ref_keep))


def test_full_size_cfg3_four_channel_1024(engine, workload_factory):
    """BASELINE config 3 at its full batch (1024 clips, 4 channels, magnitude + phase: 5.3 GB of
    features): determinism, mask / label invariants on every clip, parity on a few clips."""
    import torch
    from challenge_b200 import _lib as L
    w = workload_factory(4, seed=20203, n_bg=2, n_voice=16, n_noise=4)
    B = 1024
    d = _draw(w, B, 626, seed=3030)
    engine.upload_plan(d)
    frame, _, keep = engine.labels()
    x = engine.features(L.FEAT_MAGPHASE)
    assert tuple(x.shape) == (B, 257, 626, 8)
    y = engine.features(L.FEAT_MAGPHASE)
    assert torch.equal(x, y), 'not deterministic'
    del y
    assert bool(torch.isfinite(x).all())
    assert bool((x[..., :4] >= 0).all())                                   # magnitudes
    assert bool((x[..., 4:].abs() <= 3.1415928).all())                      # phases in [-pi, pi]
    for b in range(0, B, 97):                                               # masked cells: |.| == 0 exactly
        for size, off in d.time_masks[b]:
            assert bool((x[b, :, off:off + size, :4] == 0).all())
        for size, off in d.freq_masks[b]:
            assert bool((x[b, off:off + size, :, :4] == 0).all())
    fr = frame.cpu().numpy()
    kp = keep.cpu().numpy()
    assert set(np.unique(fr)) <= {0.0, 1.0} and np.all(kp[:, 0] == 1)
    clips = sorted(set([0, 1023] + np.random.default_rng(8).choice(B, 32, replace=False).tolist()))
    ref, ref_y, _, ref_keep = _oracle(w, d, mode='magphase', clips=clips)
    got = x[clips].cpu().numpy()
    assert nmax_err(got[..., :4], ref[..., :4]) < TOL
    for i, b in enumerate(clips):
        assert nmax_err(got[i, ..., :4], ref[i, ..., :4]) < TOL, b
    _check_phase(ref[..., :4], got[..., 4:], ref[..., 4:])
    assert np.array_equal(fr[clips], ref_y)
    assert np.array_equal(kp[clips], np.stack(ref_keep))


def test_full_size_cfg4_shards_equal_the_whole_batch(engine, workload_factory):
    """BASELINE config 4 (batch 8192 sharded over 8 GPUs): every rank receives a contiguous slice of
    the host plan.  Run here on one GPU: the 8 slices of 1024 clips must reproduce
    the whole-batch launch bit for bit -- features, labels, keep flags -- and their metric counts
    must add up to the whole batch's (what the NCCL sum all-reduce computes)."""
    import torch
    from challenge_b200 import _lib as L
    w = workload_factory(2)
    B, G = 8192, 8
    d = _draw(w, B, 626, seed=4040)
    engine.upload_plan(d)
    frame, _, keep = engine.labels()
    x = engine.features(L.FEAT_LOGMEL_MINMAX)
    yp = torch.clamp(frame + 0.35 * torch.randn(frame.shape, device=frame.device,
                                                generator=torch.Generator(frame.device).manual_seed(5)), 0, 1)
    triples, tpfpfn, _ = engine.metric_counts(frame, yp)
    total = torch.zeros(3, dtype=torch.int64, device=frame.device)
    for r in range(G):
        lo, hi = r * B // G, (r + 1) * B // G
        engine.upload_plan(d.slice(lo, hi))
        f_r, _, k_r = engine.labels()
        x_r = engine.features(L.FEAT_LOGMEL_MINMAX)
        assert torch.equal(x_r, x[lo:hi]), r
        assert torch.equal(f_r, frame[lo:hi]) and torch.equal(k_r, keep[lo:hi]), r
        t_r, c_r, _ = engine.metric_counts(f_r, yp[lo:hi])
        assert torch.equal(t_r, triples[lo:hi]), r
        total += c_r
    assert torch.equal(total, tpfpfn)
    clips = sorted(np.random.default_rng(9).choice(B, 32, replace=False).tolist())   # parity on 32 random clips
    ref, ref_y, _, ref_keep = _oracle(w, d, mode='logmel_minmax', clips=clips)
    got = x[clips].cpu().numpy()
    for i, b in enumerate(clips):
        assert nmax_err(got[i], ref[i]) < TOL, b
    assert np.array_equal(frame[clips].cpu().numpy(), ref_y)
    assert np.array_equal(keep[clips].cpu().numpy(), np.stack(ref_keep))
    xn = x[::1024].cpu().numpy()                       # every clip spans [log 1e-8, log(1 + 1e-8)]
    assert np.allclose(xn.reshape(len(xn), -1).min(1), np.log(np.float32(1e-8)), rtol=0, atol=1e-5)
    assert np.allclose(xn.reshape(len(xn), -1).max(1), 0, atol=1e-6)


@pytest.mark.parametrize('case', range(10))
def test_random_small_shapes_against_the_oracle(case):
    """Randomised shapes around the edges of the tile / claim machinery: 1..4 channels, n_frame from
    a handful to a few hundred (ragged last tile, fewer tiles than CTAs, fewer tiles than one claim),
    batch 1..9, with / without voices, noises and masks, every feature mode.  Labels, keep flags and
    masks bit-exact, features within the bars above."""
    from challenge_b200 import _lib as L
    from challenge_b200.engine import Engine
    from challenge_b200.plan import draw_batch
    from challenge_b200.synth import synthetic_banks
    from oracle import chain
    rng = np.random.default_rng(9000 + case)
    C = int(rng.integers(1, 5))
    T = int(rng.choice([3, 8, 9, 17, 40, 63, 64, 65, 130, 301]))
    B = int(rng.integers(1, 10))
    V = int(rng.integers(0, 5))
    M = int(rng.integers(0, 4)) if V > 0 else 0
    masks = bool(rng.integers(0, 2))
    bgs, voices, labels, noises = synthetic_banks(500 + case, C, n_bg=3, n_voice=7, n_noise=3,
                                                  bg_seconds=float(rng.choice([1.0, 2.5, 6.0])), lo_s=0.2, hi_s=1.5)
    eng = Engine(0)
    try:
        eng.set_mel(80)
        bf = eng.register_bank(L.BANK_BG, bgs)
        vf = eng.register_bank(L.BANK_VOICE, voices, labels=labels)
        nf = eng.register_bank(L.BANK_NOISE, noises)
        try:
            d = draw_batch(rng, B, T, bf, vf if V else None, nf if M else None, max_voices=V, max_noises=M, snr=-20,
                           min_ratio=2 / 3, n_time_masks=6 if masks and T > 24 else 0, n_freq_masks=1 if masks else 0)
        except ValueError:      # an empty offset range for this (voice length, n_frame): the reference raises too
            pytest.skip('draw raises like the reference')
        eng.upload_plan(d)
        ob, ov, on = chain.OracleBank(bgs), chain.OracleBank(voices), chain.OracleBank(noises)
        if V:
            frame, _, keep = eng.labels()
        for mode, name in ((L.FEAT_LOGMEL_MINMAX, 'logmel_minmax'), (L.FEAT_COMPLEX, 'complex'),
                           (L.FEAT_MAGPHASE, 'magphase'), (L.FEAT_MEL, 'mel')):
            got = eng.features(mode).cpu().numpy()
            ref, ref_y, _, ref_keep = chain.dataset_batch(ob, ov if V else None, labels if V else None,
                                                          on if M else None, d, mode=name)
            assert got.shape == ref.shape, (name, got.shape, ref.shape)
            if name == 'magphase':
                assert nmax_err(got[..., :C], ref[..., :C]) < TOL, name
                assert phase_report(ref[..., :C], got[..., C:], ref[..., C:])['gate1e-3'] < PHASE_TOL
            else:
                assert nmax_err(got, ref) < TOL, (name, C, T, B, V, M)
        if V:
            assert np.array_equal(frame.cpu().numpy(), ref_y)
            assert np.array_equal(keep.cpu().numpy(), np.stack(ref_keep))
    finally:
        eng.close()
