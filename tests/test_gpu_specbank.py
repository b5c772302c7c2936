"""Spectrogram-format banks (the reference's own pickled ``[257, t, 2C]`` lists, utils.py:88-94;
SURVEY.md 8f rank 2) through make_pipeline / merge_complex_specs / the engine: the reference's
pipeline tests restated (pipeline_test.py:13-74 use random spectrograms), and parity of every
feature mode with the CPU oracle, which mixes spectrograms exactly like pipeline.py."""
import numpy as np
import pytest

from conftest import nmax_err, phase_err

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope='module')
def mods(engine):
    import challenge_b200
    from challenge_b200 import data_utils, engine as E, metrics, pipeline, transforms
    E._engines[0] = engine
    challenge_b200.set_seed(0)
    return pipeline, transforms, data_utils, metrics


def _spec_banks(seed, chan2=4, n_bg=5, n_voice=12, n_noise=6, n_classes=3, bg_t=(20, 200),
                v_t=(3, 60), n_t=(3, 60)):
    rng = np.random.default_rng(seed)
    F = 257

    def spec(t, tail_zero=False):
        x = rng.standard_normal((F, t, chan2)).astype(np.float32)
        if tail_zero:       # trailing silent frames exercise the `> 0` activity mask (pipeline_test.py:21-24)
            x[:, t - int(rng.integers(0, max(t // 3, 1))):] = 0
        return x
    bgs = [spec(int(rng.integers(*bg_t))) for _ in range(n_bg)]
    voices = [spec(int(rng.integers(*v_t)), True) for _ in range(n_voice)]
    labels = np.eye(n_classes, dtype=np.float32)[rng.integers(0, n_classes, n_voice)]
    noises = [spec(int(rng.integers(*n_t))) for _ in range(n_noise)]
    return bgs, voices, labels, noises


# ---- pipeline_test.py restated against the drop-in ----
def test_merge_complex_specs_reference_test(mods):
    P = mods[0]
    rng = np.random.default_rng(0)
    freq, chan, n_classes, n_frame = 257, 4, 30, 10
    background = rng.standard_normal((freq, 8, chan)).astype(np.float32)
    n_voices = 4
    voices = rng.standard_normal((n_voices, freq, n_frame, chan)).astype(np.float32)
    lens = rng.integers(1, n_frame, size=n_voices)
    voices *= (np.arange(n_frame)[None, :] < lens[:, None]).reshape(n_voices, 1, n_frame, 1)
    labels = np.eye(n_classes, dtype=np.float32)[rng.integers(1, n_frame, size=n_voices)]
    n_noises = 2
    noises = rng.standard_normal((n_noises, freq, n_frame, chan)).astype(np.float32)
    lens = rng.integers(1, n_frame, size=n_noises)
    noises *= (np.arange(n_frame)[None, :] < lens[:, None]).reshape(n_noises, 1, n_frame, 1)
    spec, l = P.merge_complex_specs(background, (voices, labels), noises, n_frame=n_frame,
                                    n_classes=n_classes)
    assert tuple(spec.shape) == (freq, n_frame, chan)
    assert tuple(l.shape) == (n_voices, n_frame, n_classes)


def test_make_pipeline_reference_test(mods):
    P = mods[0]
    rng = np.random.default_rng(1)
    freq, chan, n_classes, n_frame = 257, 4, 30, 30
    backgrounds = [rng.standard_normal((freq, int(rng.integers(1, n_frame * 2)), chan)) for _ in range(30)]
    voices = [rng.standard_normal((freq, int(rng.integers(1, n_frame // 2)), chan)) for _ in range(40)]
    labels = np.eye(n_classes, dtype=np.float32)[rng.integers(n_classes, size=(40,))]
    noises = [rng.standard_normal((freq, int(rng.integers(1, n_frame // 2)), chan)) for _ in range(50)]
    pipeline = P.make_pipeline(backgrounds, voices, labels, noises, n_frame=n_frame, max_voices=4,
                               max_noises=4, n_classes=n_classes)
    n = 0
    for s, l in pipeline.take(3):
        assert tuple(s.shape) == (freq, n_frame, chan)
        assert tuple(l.shape) == (4, n_frame, n_classes)
        n += 1
    assert n == 3


# ---- values: every mode against the oracle (which mixes spectrograms like pipeline.py) ----
def _run(engine, banks, d, mode_name, **plan_kw):
    from challenge_b200 import _lib as L
    from oracle import chain
    bgs, voices, labels, noises = banks
    modes = {'complex': L.FEAT_COMPLEX, 'magphase': L.FEAT_MAGPHASE, 'log_magphase': L.FEAT_LOG_MAGPHASE,
             'mel': L.FEAT_MEL, 'logmel': L.FEAT_LOGMEL, 'logmel_minmax': L.FEAT_LOGMEL_MINMAX}
    engine.upload_plan(d, **plan_kw)
    frame, vtk, keep = engine.labels(want_vtk=True)
    x = engine.features(modes[mode_name])
    remap = {L.REMAP_NONE: None, L.REMAP_STEREO_MONO: 'stereo_mono', L.REMAP_MERGE_AUG: 'merge_aug'}[
        plan_kw.get('chan_remap', L.REMAP_NONE)]
    ref, ref_y, ref_vtk, ref_keep = chain.dataset_batch(
        bgs, voices, labels, noises, d, n_classes=labels.shape[1], mode=mode_name, remap=remap,
        n_out_chan=plan_kw.get('n_out_chan', 0), stft_filter=plan_kw.get('stft_filter', 0))
    return _np(x), ref, _np(frame), ref_y, _np(vtk), ref_vtk, _np(keep), ref_keep


@pytest.mark.parametrize('chan2', [4, 8, 2])
def test_specbank_complex_is_bit_exact(engine, chan2):
    from challenge_b200 import _lib as L
    from challenge_b200.plan import draw_batch
    banks = _spec_banks(10 + chan2, chan2=chan2)
    bgs, voices, labels, noises = banks
    bf = engine.register_bank(L.BANK_BG, bgs)
    vf = engine.register_bank(L.BANK_VOICE, voices, labels=labels)
    nf = engine.register_bank(L.BANK_NOISE, noises)
    assert bf.tolist() == [x.shape[1] for x in bgs] and vf.tolist() == [x.shape[1] for x in voices]
    d = draw_batch(np.random.default_rng(3), 6, 80, bf, vf, nf, max_voices=5, max_noises=3,
                   min_ratio=2 / 3, n_time_masks=6, n_freq_masks=1)
    x, ref, frame, ref_y, vtk, ref_vtk, keep, ref_keep = _run(engine, banks, d, 'complex')
    assert x.shape == ref.shape == (6, 257, 80, chan2)
    assert np.array_equal(x, ref)                       # same products and sums in the same order
    assert np.array_equal(frame, ref_y)
    assert np.array_equal(vtk, np.stack(ref_vtk))
    assert np.array_equal(keep, np.stack(ref_keep))
    # voice activity is `any coefficient > 0` of the stored spectrogram (pipeline.py:55)
    act = np.zeros(voices[0].shape[1], np.uint8)
    engine.lib.iris_bank_activity(engine._ctx, 0, act.ctypes.data)
    assert np.array_equal(act, (voices[0].max(axis=(0, 2)) > 0).astype(np.uint8))


@pytest.mark.parametrize('mode', ['magphase', 'log_magphase', 'mel', 'logmel', 'logmel_minmax'])
def test_specbank_feature_modes(engine, mode):
    from challenge_b200 import _lib as L
    from challenge_b200.plan import draw_batch
    engine.set_mel(80)
    banks = _spec_banks(21)
    bgs, voices, labels, noises = banks
    bf = engine.register_bank(L.BANK_BG, bgs)
    vf = engine.register_bank(L.BANK_VOICE, voices, labels=labels)
    nf = engine.register_bank(L.BANK_NOISE, noises)
    d = draw_batch(np.random.default_rng(4), 5, 100, bf, vf, nf, max_voices=4, max_noises=3,
                   min_ratio=2 / 3, n_time_masks=6, n_freq_masks=1)
    x, ref, frame, ref_y, *_ = _run(engine, banks, d, mode)
    assert x.shape == ref.shape
    assert np.array_equal(frame, ref_y)
    if mode in ('magphase', 'log_magphase'):
        C = x.shape[-1] // 2
        mag_ref = np.sqrt(np.square(_run(engine, banks, d, 'complex')[1]).reshape(5, 257, 100, 2, C).sum(3))
        if mode == 'magphase':
            assert nmax_err(x[..., :C], ref[..., :C]) <= 1e-6
        else:       # log of tiny magnitudes: gate like the phase
            sel = mag_ref > 1e-3 * mag_ref.max()
            assert np.abs(x[..., :C][sel] - ref[..., :C][sel]).max() <= 1e-4
            assert np.array_equal(x[..., :C][mag_ref == 0], ref[..., :C][mag_ref == 0])   # log(1e-8) on masked cells
        assert phase_err(mag_ref, x[..., C:], ref[..., C:]) <= 1e-5
    else:
        assert nmax_err(x, ref) <= 1e-4


@pytest.mark.parametrize('remap', ['stereo_mono', 'merge_aug'])
def test_specbank_remap_and_filter(engine, remap):
    from challenge_b200 import _lib as L
    from challenge_b200.plan import draw_batch
    banks = _spec_banks(33)
    bgs, voices, labels, noises = banks
    bf = engine.register_bank(L.BANK_BG, bgs)
    vf = engine.register_bank(L.BANK_VOICE, voices, labels=labels)
    nf = engine.register_bank(L.BANK_NOISE, noises)
    n_out = 3 if remap == 'stereo_mono' else 5
    d = draw_batch(np.random.default_rng(5), 4, 64, bf, vf, nf, max_voices=4, max_noises=3,
                   n_time_masks=6, n_freq_masks=1, merge_extra=n_out - 2 if remap == 'merge_aug' else 0)
    kw = dict(stft_filter=16, chan_remap=L.REMAP_STEREO_MONO if remap == 'stereo_mono' else L.REMAP_MERGE_AUG,
              n_out_chan=n_out)
    x, ref, *_ = _run(engine, banks, d, 'complex', **kw)
    assert x.shape == ref.shape == (4, 257, 64, 2 * n_out)
    assert nmax_err(x, ref) <= 1e-6
    assert np.all(x[:, 1:17] == 0)


def test_specbank_edge_cases(engine):
    """One-frame items, a background shorter than the clip (tiled, pipeline.py:29-35), no noise
    stream, max_voices = 1, and the reference's empty-range error."""
    from challenge_b200 import _lib as L
    from challenge_b200.errors import InvalidArgumentError
    from challenge_b200.plan import draw_batch
    banks = _spec_banks(44, bg_t=(1, 4), v_t=(1, 3), n_t=(1, 2), n_bg=3, n_voice=5, n_noise=2)
    bgs, voices, labels, noises = banks
    bf = engine.register_bank(L.BANK_BG, bgs)
    vf = engine.register_bank(L.BANK_VOICE, voices, labels=labels)
    nf = engine.register_bank(L.BANK_NOISE, noises)
    d = draw_batch(np.random.default_rng(6), 8, 12, bf, vf, nf, max_voices=3, max_noises=2)
    x, ref, frame, ref_y, vtk, ref_vtk, keep, ref_keep = _run(engine, banks, d, 'complex')
    assert np.array_equal(x, ref) and np.array_equal(frame, ref_y) and np.array_equal(keep, np.stack(ref_keep))
    d = draw_batch(np.random.default_rng(7), 3, 9, bf, vf, None, max_voices=1, max_noises=0)
    x, ref, frame, ref_y, *_ = _run(engine, (bgs, voices, labels, None), d, 'complex')
    assert np.array_equal(x, ref) and np.array_equal(frame, ref_y)
    # a voice group as long as the clip with min_ratio = 1 leaves no offset to draw (pipeline.py:68-69)
    long_voices = [np.ones((257, 12, 4), np.float32)] * 2
    engine.register_bank(L.BANK_VOICE, long_voices, labels=np.eye(3, dtype=np.float32)[[0, 1]])
    with pytest.raises((InvalidArgumentError, ValueError)):
        draw_batch(np.random.default_rng(8), 2, 12, bf, np.array([12, 12]), None, max_voices=2, min_ratio=1)
    # mixing bank formats is refused
    from challenge_b200.synth import synthetic_banks
    wb, *_ = synthetic_banks(1, 2, n_bg=1, n_voice=1, n_noise=1, bg_seconds=1.0)
    engine.register_bank(L.BANK_BG, wb)
    d = draw_batch(np.random.default_rng(9), 1, 12, engine.bank_frames[L.BANK_BG], np.array([12, 12]), None,
                   max_voices=2, min_ratio=2 / 3)
    with pytest.raises(ValueError):
        engine.upload_plan(d)


def test_specbank_agrees_with_waveform_path(engine):
    """The two formats of the same audio: waveform banks through the fused FFT kernel, and the
    spectrograms load_wav makes of them through the spectrogram-domain mix -- same draws, same
    features (the linearity identity of DESIGN.md 2, checked GPU against GPU)."""
    from challenge_b200 import _lib as L
    from challenge_b200.plan import draw_batch
    from challenge_b200.synth import synthetic_banks
    engine.set_mel(80)
    bgs, voices, labels, noises = synthetic_banks(5, 2, n_bg=3, n_voice=8, n_noise=4, bg_seconds=4.0)
    bf = engine.register_bank(L.BANK_BG, bgs)
    vf = engine.register_bank(L.BANK_VOICE, voices, labels=labels)
    nf = engine.register_bank(L.BANK_NOISE, noises)
    d = draw_batch(np.random.default_rng(12), 6, 200, bf, vf, nf, max_voices=4, max_noises=2,
                   n_time_masks=6, n_freq_masks=1)
    engine.upload_plan(d)
    frame_w, _, keep_w = engine.labels()
    out_w = {m: _np(engine.features(m)) for m in (L.FEAT_COMPLEX, L.FEAT_LOGMEL_MINMAX)}
    frame_w, keep_w = _np(frame_w), _np(keep_w)
    sb = [_np(engine.stft(w, normalize=True)) for w in bgs]
    sv = [_np(engine.stft(w, normalize=True)) for w in voices]
    sn = [_np(engine.stft(w, normalize=True)) for w in noises]
    assert engine.register_bank(L.BANK_BG, sb).tolist() == bf.tolist()
    engine.register_bank(L.BANK_VOICE, sv, labels=labels)
    engine.register_bank(L.BANK_NOISE, sn)
    engine.upload_plan(d)
    frame_s, _, keep_s = engine.labels()
    assert np.array_equal(_np(frame_s), frame_w) and np.array_equal(_np(keep_s), keep_w)
    for m, tol in ((L.FEAT_COMPLEX, 2e-6), (L.FEAT_LOGMEL_MINMAX, 1e-4)):
        assert nmax_err(_np(engine.features(m)), out_w[m]) <= tol


@pytest.mark.parametrize('fmt', ['spec', 'wave'])
def test_seperate_noise_voice(engine, mods, fmt):
    """merge_complex_specs(seperate_noise_voice=True) (pipeline.py:37-38, 82-83, 104-108): the
    label becomes (label, only_voice, only_noise), in both bank formats; and the 'se' chain of
    sj_train.make_dataset (sj_train.py:99-105) on top of it."""
    from challenge_b200 import _lib as L
    from challenge_b200.plan import draw_batch
    from challenge_b200.synth import synthetic_banks
    from oracle import chain, pipeline as OP
    P, _, DU, _ = mods
    if fmt == 'spec':
        bgs, voices, labels, noises = _spec_banks(55)
        o_b, o_v, o_n = bgs, voices, noises
    else:
        bgs, voices, labels, noises = synthetic_banks(6, 2, n_bg=3, n_voice=8, n_noise=4, bg_seconds=3.0)
        o_b, o_v, o_n = chain.OracleBank(bgs), chain.OracleBank(voices), chain.OracleBank(noises)
    bf = engine.register_bank(L.BANK_BG, bgs)
    vf = engine.register_bank(L.BANK_VOICE, voices, labels=labels)
    nf = engine.register_bank(L.BANK_NOISE, noises)
    d = draw_batch(np.random.default_rng(13), 5, 90, bf, vf, nf, max_voices=4, max_noises=3)
    engine.upload_plan(d)
    engine.labels()
    spec = _np(engine.features(L.FEAT_COMPLEX))
    only_voice = _np(engine.features(L.FEAT_COMPLEX, select=L.SELECT_VOICES))
    only_noise = _np(engine.features(L.FEAT_COMPLEX, select=L.SELECT_BG_NOISE))
    for b in range(5):
        ref_spec, (ref_l, ref_v, ref_n) = chain.synth_clip(o_b, o_v, labels, o_n, d, b, seperate_noise_voice=True)
        if fmt == 'spec':
            assert np.array_equal(spec[b], ref_spec)
            assert np.array_equal(only_voice[b], ref_v)
            assert np.array_equal(only_noise[b], ref_n)
        else:
            assert nmax_err(spec[b], ref_spec) <= 1e-5
            assert np.abs(only_voice[b] - ref_v).max() <= 1e-5 * np.abs(ref_spec).max()
            assert nmax_err(only_noise[b], ref_n) <= 1e-5
    # the dataset form, through the 'se' chain: speech_enhancement_preprocess -> batch -> label_downsample(32)
    ds = P.make_pipeline(bgs, voices, labels, noises, n_frame=64, max_voices=4, max_noises=3,
                         seperate_noise_voice=True)
    ds = ds.map(DU.speech_enhancement_preprocess).batch(3).map(DU.label_downsample(32))
    n = 0
    for x, y in ds.take(2):
        C = 2
        assert tuple(x.shape) == (3, 256, 64, C)
        assert isinstance(y, tuple) and len(y) == 3
        assert tuple(y[0].shape) == (3, 2, 3)                 # ceil(64 / 32) pooled frame labels
        # reference quirk kept (data_utils.py:147): y[1], y[2] are cut to x.shape[-1] // 2 AFTER x was
        # halved, i.e. to chan // 2 channels
        assert tuple(y[1].shape) == (3, 256, 64, C // 2) and tuple(y[2].shape) == (3, 256, 64, C // 2)
        n += 1
    assert n == 2
