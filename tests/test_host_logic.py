"""CPU tests of the host side: draws, placement arithmetic, the C-ABI library (loads,
exports every symbol include/iris.h declares, fails loudly without a GPU), the FFT core's
math compiled for the host, and the product never importing the oracle."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_draw_ranges_follow_reference():
    from challenge_b200.plan import draw_batch, placement
    rng = np.random.default_rng(0)
    bf = np.array([626, 300, 40])
    vf = rng.integers(32, 251, 50)
    nf = rng.integers(32, 251, 20)
    T, V, M = 626, 7, 2
    d = draw_batch(rng, 500, T, bf, vf, nf, V, M, snr=-20, min_ratio=1, n_time_masks=6,
                   n_freq_masks=1, merge_extra=2)
    bgT = bf[d.bg_id]
    tiled = bgT * ((T + bgT - 1) // bgT)
    assert np.all((d.bg_offset >= 0) & (d.bg_offset <= tiled - T))          # pipeline.py:35
    assert np.all((d.n_voices >= 1) & (d.n_voices < V))                      # :43
    assert np.all((d.n_noises >= 0) & (d.n_noises < M))                      # :87
    assert np.all((d.voice_u >= 0) & (d.voice_u < 2)) and d.voice_u.dtype == np.float32
    assert np.array_equal(d.voice_gain, np.power(np.float32(10), -d.voice_u, dtype=np.float32))
    for b in range(500):
        vP = vf[d.voice_id[b]].max()
        pad, length = placement(T, vP, 1)
        assert pad == T - vP and length == vP + 2 * pad
        assert np.all(d.voice_offset[b, :d.n_voices[b]] < length - T)         # :69
        nP = nf[d.noise_id[b]].max()
        _, nlen = placement(T, nP, 0.5)
        assert np.all(d.noise_offset[b, :d.n_noises[b]] <= nlen - T)          # :103
    size, off = d.time_masks[..., 0], d.time_masks[..., 1]
    assert np.all((size >= 0) & (size < 24) & (off >= 0) & (off < T - size))  # transforms.py:25-26
    size, off = d.freq_masks[..., 0], d.freq_masks[..., 1]
    assert np.all((size >= 0) & (size < 16) & (off >= 0) & (off < 257 - size))
    assert np.all((d.merge_factor >= 0.1) & (d.merge_factor < 0.9))           # data_utils.py:109
    s = d.slice(100, 164)
    assert s.batch == 64 and np.array_equal(s.voice_id, d.voice_id[100:164])


def test_placement_truncates_in_float32():
    from challenge_b200.plan import placement
    # 2/3 * 100 = 66.67 -> int32 66 ; pad = 34
    assert placement(100, 100, 2 / 3) == (34, 168)
    assert placement(626, 626, 1) == (0, 626)
    assert placement(10, 30, 0.5) == (0, 30)


def test_empty_offset_range_raises_in_planner():
    from challenge_b200.errors import InvalidArgumentError
    from challenge_b200.plan import draw_batch
    with pytest.raises(InvalidArgumentError):
        draw_batch(np.random.default_rng(0), 2, 100, np.array([100]), np.array([100, 50]),
                   None, max_voices=2, min_ratio=1)


def test_shuffle_stream_visits_everything():
    from challenge_b200.plan import ShuffleStream
    s = ShuffleStream(10, np.random.default_rng(0))
    ids = s.take(1000)
    assert set(ids.tolist()) == set(range(10))
    # repeat().shuffle(n): over a long run every id appears about equally often
    assert np.bincount(ids, minlength=10).min() > 60


def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'iris.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(iris_[a-z0-9_]+)\s*\(', text)))


def test_library_loads_and_exports_every_declared_symbol():
    from challenge_b200 import _lib
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 26
    assert sorted(_lib.SIGNATURES) == syms, 'include/iris.h and the ctypes binding differ'
    for name in syms:
        assert hasattr(lib, name), 'libiris.so does not export %s' % name
        assert name in _lib.SIGNATURES, 'no ctypes signature for %s' % name
    assert lib.iris_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from challenge_b200 import _lib
    lib = _lib.load()
    ctx = ctypes.c_void_p()
    rc = lib.iris_ctx_create(0, ctypes.byref(ctx))
    assert rc == _lib.IRIS_ERR_CUDA and b'no CPU fallback' in lib.iris_last_error()
    from challenge_b200.engine import Engine
    with pytest.raises(Exception):
        Engine(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'challenge_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f


def test_fftcore_math_on_host(tmp_path):
    """fftcore.cuh compiled with g++: the 16-lane x 32-point decomposition, the 4-round
    exchange geometry and the lane-local two-channel split reproduce a float64 DFT."""
    exe = str(tmp_path / 'fftcore_host')
    subprocess.check_call(['g++', '-std=c++17', '-O2', '-x', 'c++', '-I',
                           os.path.join(ROOT, 'challenge_b200', 'csrc'),
                           os.path.join(ROOT, 'tests', 'host', 'fftcore_host.cpp'), '-o', exe])
    r = json.loads(subprocess.check_output([exe]).decode())
    assert r['bad_seen'] == 0
    assert r['err_small'] < 2e-6
    assert r['err512'] / r['max_abs'] < 1e-6
    assert r['im_dc'] == 0.0 and r['im_nyq'] == 0.0


def test_fftwarp_math_on_host(tmp_path):
    """fftwarp.cuh (the warp-per-frame FFT of the fused kernel) compiled with g++: 32 emulated
    lanes, the exchange offsets, the DIF + partner/mirror maps and the two-channel split
    reproduce a float64 windowed DFT; bins < 128 need only half of pass 2's outputs."""
    exe = str(tmp_path / 'fftwarp_host')
    subprocess.check_call(['g++', '-std=c++17', '-O2', '-x', 'c++', '-I',
                           os.path.join(ROOT, 'challenge_b200', 'csrc'),
                           os.path.join(ROOT, 'tests', 'host', 'fftwarp_host.cpp'), '-o', exe])
    r = json.loads(subprocess.check_output([exe]).decode())
    assert r['bad_map'] == 0 and r['missing'] == 0 and r['bad_prune'] == 0
    assert r['err512'] / r['max_abs'] < 1e-6
    assert r['im_dc'] == 0.0 and r['im_nyq'] == 0.0


def test_default_mel_matrix_is_the_tf_restatement():
    from challenge_b200.engine import default_mel_matrix
    from oracle.transforms import linear_to_mel_weight_matrix
    assert np.array_equal(default_mel_matrix(80), linear_to_mel_weight_matrix(80, 257, 16000))
    assert np.array_equal(default_mel_matrix(40, lower_edge_hertz=80.0, upper_edge_hertz=7600.0),
                          linear_to_mel_weight_matrix(40, 257, 16000, 80.0, 7600.0))


def test_dropin_modules_mirror_the_reference_surface():
    """Every public function of the reference's hot-path modules exists under the same name
    (SURVEY.md 8b); importing the modules needs neither a GPU nor the oracle."""
    from challenge_b200 import data_utils, metrics, pipeline, transforms
    want = {
        pipeline: ['merge_complex_specs', 'make_pipeline'],
        transforms: ['mask', 'random_shift', 'magphase_to_mel', 'log_magphase',
                     'minmax_norm_magphase', 'complex_to_magphase', 'magphase_to_complex',
                     'phase_vocoder', 'EPSILON'],
        data_utils: ['load_wav', 'normalize', 'minmax', 'log_on_mel', 'augment', 'to_frame_labels',
                     'mono_chan', 'stereo_mono', 'label_downsample', 'random_merge_aug',
                     'multiply_label', 'stft_filter', 'speech_enhancement_preprocess'],
        metrics: ['er_score', 'f1_score', 'cos_sim'],
    }
    for mod, names in want.items():
        for n in names:
            assert hasattr(mod, n), '%s.%s missing' % (mod.__name__, n)


def test_dataset_lowering_recognises_the_sj_train_chain():
    """sj_train.make_dataset's stage order (sj_train.py:107-123) lowers to one fused launch."""
    from challenge_b200 import _lib as L
    from challenge_b200 import data_utils as D, transforms as TR
    from challenge_b200.pipeline import IrisDataset
    ds = IrisDataset({})
    ds = ds.map(D.to_frame_labels).map(D.augment).map(D.stft_filter(3)).batch(12)
    ds = ds.map(TR.complex_to_magphase).map(TR.magphase_to_mel(80)).map(D.minmax).map(D.log_on_mel)
    fused, rest = ds.prefetch(-1)._lower()
    assert rest == [] and fused['mode'] == L.FEAT_LOGMEL_MINMAX and fused['augment']
    assert fused['frame_labels'] and fused['filt'] == 3 and fused['n_mels'] == 80
    # 'nominmax' variant (sj_train.py:121) and a stage the fused kernel does not absorb
    ds = IrisDataset({}).map(D.to_frame_labels).batch(4).map(TR.complex_to_magphase)
    ds = ds.map(TR.magphase_to_mel(40)).map(D.log_on_mel).map(D.label_downsample(32))
    fused, rest = ds._lower()
    assert fused['mode'] == L.FEAT_LOGMEL and fused['n_mels'] == 40 and len(rest) == 1
    # out-of-order stages stay un-fused and keep their order
    ds = IrisDataset({}).map(D.augment).map(D.to_frame_labels).batch(2)
    fused, rest = ds._lower()
    assert fused['augment'] and not fused['frame_labels'] and len(rest) == 2


def test_resample_len_matches_the_kaldi_port_without_a_gpu():
    """iris_resample_len is host arithmetic (no device needed): it must agree with the oracle's
    restatement of kaldi.py::_get_num_LR_output_samples for every rate pair / length."""
    from challenge_b200 import _lib
    from oracle.data_utils import _lr_num_output_samples
    lib = _lib.load()
    rng = np.random.default_rng(3)
    for orig, new in ((44100, 16000), (48000, 16000), (8000, 16000), (22050, 16000), (11025, 16000),
                      (32000, 16000), (16000, 16000), (96000, 16000), (16000, 8000)):
        for n in [1, 2, 255, 256, 441, 16000, 44100, 160000] + [int(v) for v in rng.integers(1, 500000, 20)]:
            assert lib.iris_resample_len(n, orig, new) == _lr_num_output_samples(n, orig, new), (orig, new, n)
    assert lib.iris_resample_len(0, 44100, 16000) == 0


def test_numa_binding_is_a_no_op_without_nvml():
    """dist.bind_to_gpu_numa never raises: without a GPU / NVML it leaves the affinity alone."""
    from challenge_b200.dist import bind_to_gpu_numa
    before = os.sched_getaffinity(0)
    cores = bind_to_gpu_numa(0)
    assert cores is None or set(cores) <= before
    if cores is None:
        assert os.sched_getaffinity(0) == before
    else:
        os.sched_setaffinity(0, before)


def test_default_mel_matrix_has_the_shape_the_fixed_kernel_instances_assume():
    """k_fused<FM_MEL, 4, EPI_C2 | ...> unrolls the projection for filters of at most 2 / 4 / 6 bins
    in the three rounds of 32 and bins below 128 (fixed_mel_L in k_fused.cu; the launcher checks the
    same and falls back to the generic instance otherwise).  The reference's matrix
    (transforms.py:55, linear_to_mel_weight_matrix(80, 257, 16000)) has exactly that shape."""
    from oracle.transforms import linear_to_mel_weight_matrix
    W = linear_to_mel_weight_matrix(80, 257, 16000)
    lens, hi = [], 0
    for m in range(80):
        nz = np.nonzero(W[:, m])[0]
        assert np.all(np.diff(nz) == 1)                      # contiguous support
        lens.append(len(nz))
        hi = max(hi, int(nz[-1]))
    rounds = [max(lens[32 * r:32 * r + 32]) for r in range(3)]
    assert [n + (n & 1) for n in rounds] == [2, 4, 6] and hi < 128
    src = open(os.path.join(ROOT, 'challenge_b200', 'csrc', 'k_fused.cu')).read()
    assert 'return r == 0 ? 2 : (r == 1 ? 4 : (r == 2 ? 6 : 0));' in src
