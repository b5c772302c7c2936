"""Oracle restatement of the reference's ``transforms.py`` (TEST INFRASTRUCTURE ONLY)."""
import numpy as np

from . import EPSILON


def mask(specs, axis, max_mask_size=None, n_mask=1, *, draws):
    """transforms.py:12-40.  ``draws`` = ``n_mask`` pairs ``(size, offset)`` with
    ``size in [0, max_mask_size)`` (line 25) and ``offset in [0, total-size)``
    (line 26), consumed in that order.  The mask is applied by multiplication
    (line 40), so masked cells keep the sign of their zero.
    """
    specs = np.asarray(specs)
    total = specs.shape[axis]
    if max_mask_size is None:
        max_mask_size = total
    draws = list(draws)
    assert len(draws) == n_mask
    shape = [1] * specs.ndim
    shape[axis] = total
    m = np.ones(total, dtype=specs.dtype)
    for size, offset in draws:
        size, offset = int(size), int(offset)
        assert 0 <= size < max_mask_size and 0 <= offset < total - size
        piece = np.concatenate((np.ones(offset, m.dtype), np.zeros(size, m.dtype),
                                np.ones(total - size - offset, m.dtype)))
        m = m * piece
    return specs * m.reshape(shape)


def random_shift(specs, axis=0, width=16, *, offset):
    """transforms.py:43-47 -- zero-pad ``width`` both sides, crop at ``offset``
    (``offset in [0, 2*width]``, tf.image.random_crop)."""
    specs = np.asarray(specs)
    pad = [(0, 0) if i != axis else (width, width) for i in range(specs.ndim)]
    new = np.pad(specs, pad)
    sl = [slice(None)] * specs.ndim
    sl[axis] = slice(offset, offset + specs.shape[axis])
    return new[tuple(sl)]


def linear_to_mel_weight_matrix(num_mel_bins=20, num_spectrogram_bins=129,
                                sample_rate=8000, lower_edge_hertz=125.0,
                                upper_edge_hertz=3800.0):
    """``tf.signal.linear_to_mel_weight_matrix`` of TF 2.2 [TF-sem], restated
    op-for-op in float32 (called at transforms.py:55-56).  PARITY UNPINNED.

    TF 2.2's LinSpace CPU kernel is ``start + step*i`` with
    ``step = (stop-start)/(num-1)`` evaluated in T=float.
    """
    f32 = np.float32

    def linspace(start, stop, num):
        start, stop = f32(start), f32(stop)
        step = f32((stop - start) / f32(num - 1))
        return (start + step * np.arange(num, dtype=np.float32)).astype(np.float32)

    def hertz_to_mel(f):
        return (f32(1127.0) * np.log(f32(1.0) + (np.asarray(f, np.float32) / f32(700.0)))
                ).astype(np.float32)

    nyquist = f32(sample_rate) / f32(2.0)
    lin = linspace(0.0, nyquist, num_spectrogram_bins)[1:]
    spec_mel = hertz_to_mel(lin)[:, None]
    edges = linspace(hertz_to_mel(f32(lower_edge_hertz)), hertz_to_mel(f32(upper_edge_hertz)),
                     num_mel_bins + 2)
    lower = edges[None, :-2]
    center = edges[None, 1:-1]
    upper = edges[None, 2:]
    lower_slopes = (spec_mel - lower) / (center - lower)
    upper_slopes = (upper - spec_mel) / (upper - center)
    w = np.maximum(f32(0.0), np.minimum(lower_slopes, upper_slopes)).astype(np.float32)
    return np.pad(w, [[1, 0], [0, 0]])


def magphase_to_mel(num_mel_bins=80, num_spectrogram_bins=257, sample_rate=16000, **kwargs):
    """transforms.py:51-77 -- drop phase half, dense tensordot over F, transpose."""
    mel_matrix = linear_to_mel_weight_matrix(num_mel_bins, num_spectrogram_bins,
                                             sample_rate, **kwargs)

    def _magphase_to_mel(x, y=None):
        x = np.asarray(x, np.float32)
        x = x[..., :x.shape[-1] // 2]
        x = np.tensordot(x, mel_matrix, axes=[-3, 0]).astype(np.float32)
        if x.ndim == 4:
            x = x.transpose(0, 3, 1, 2)
        elif x.ndim == 3:
            x = x.transpose(2, 0, 1)
        else:
            raise ValueError('len(x.shape) must be 3 or 4')
        x = np.ascontiguousarray(x)
        if y is None:
            return x
        return x, y
    return _magphase_to_mel


def log_magphase(specs, labels=None, n_chan=2):
    """transforms.py:80-86."""
    specs = np.asarray(specs)
    specs = np.concatenate([np.log(specs[..., :n_chan] + EPSILON), specs[..., n_chan:]],
                           axis=-1)
    if labels is not None:
        return specs, labels
    return specs


def minmax_norm_magphase(specs, labels=None):
    """transforms.py:89-107 -- note ``+EPSILON`` in the denominator, no safe_div."""
    specs = np.asarray(specs)
    n_chan = specs.shape[-1] // 2
    mag = specs[..., :n_chan]
    phase = specs[..., n_chan:]
    axis = tuple(range(1, specs.ndim))
    mag_max = mag.max(axis=axis, keepdims=True)
    mag_min = mag.min(axis=axis, keepdims=True)
    phase_max = phase.max(axis=axis, keepdims=True)
    phase_min = phase.min(axis=axis, keepdims=True)
    eps = specs.dtype.type(EPSILON)
    specs = np.concatenate([(mag - mag_min) / (mag_max - mag_min + eps),
                            (phase - phase_min) / (phase_max - phase_min + eps)], axis=-1)
    if labels is not None:
        return specs, labels
    return specs


def complex_to_magphase(complex_tensor, y=None):
    """transforms.py:111-123."""
    complex_tensor = np.asarray(complex_tensor)
    n_chan = complex_tensor.shape[-1] // 2
    real = complex_tensor[..., :n_chan]
    img = complex_tensor[..., n_chan:]
    mag = np.sqrt(real ** 2 + img ** 2)
    phase = np.arctan2(img, real)
    magphase = np.concatenate([mag, phase], axis=-1)
    if y is None:
        return magphase
    return magphase, y


def magphase_to_complex(magphase):
    """transforms.py:126-134."""
    magphase = np.asarray(magphase)
    n_chan = magphase.shape[-1] // 2
    mag = magphase[..., :n_chan]
    phase = magphase[..., n_chan:]
    return np.concatenate([mag * np.cos(phase), mag * np.sin(phase)], axis=-1)


def phase_vocoder(complex_spec, rate=1.):
    """transforms.py:137-195 (float32)."""
    complex_spec = np.asarray(complex_spec, np.float32)
    if rate == 1:
        return complex_spec
    f32 = np.float32
    freq = complex_spec.shape[0]
    hop_length = freq - 1
    n_chan = complex_spec.shape[-1] // 2

    def angle(spec):
        return np.arctan2(spec[..., n_chan:], spec[..., :n_chan])

    # tf.linspace on float32 arguments [TF-sem]: start + step * i in float32 with
    # step = (stop - start) / (num - 1); stop = float32(pi) * hop
    stop = f32(np.pi) * f32(hop_length)
    step = f32(stop / f32(freq - 1))
    phase_advance = (step * np.arange(freq, dtype=np.float32)).astype(np.float32).reshape(-1, 1, 1)
    n_t = complex_spec.shape[1]
    # tf.range(0, T, rate, dtype=float32): ceil(T/rate) elements, start + i*delta
    n_steps = int(np.ceil(n_t / rate))
    time_steps = (f32(0) + np.arange(n_steps, dtype=np.float32) * f32(rate)).astype(np.float32)

    spec = np.pad(complex_spec, [(0, 0), (0, 2), (0, 0)])
    idx0 = time_steps.astype(np.int32)
    idx1 = (time_steps + f32(1)).astype(np.int32)
    spec_0 = spec[:, idx0]
    spec_1 = spec[:, idx1]
    angle_0 = angle(spec_0)
    angle_1 = angle(spec_1)

    def norm(s):
        s = s.reshape(freq, -1, 2, n_chan).transpose(0, 1, 3, 2)
        return np.sqrt(np.sum(s * s, axis=-1, dtype=np.float32))

    norm_0 = norm(spec_0)
    norm_1 = norm(spec_1)
    phase_0 = angle(spec[..., :1, :])
    phase = angle_1 - angle_0 - phase_advance
    two_pi = f32(2 * np.pi)
    phase = phase - two_pi * np.round(phase / two_pi)
    phase = phase + phase_advance
    phase = np.concatenate([phase_0, phase[:, :-1]], axis=1)
    phase_acc = np.cumsum(phase, 1, dtype=np.float32)
    alphas = np.mod(time_steps, f32(1.)).reshape(1, -1, 1)
    mag = alphas * norm_1 + (f32(1) - alphas) * norm_0
    real = mag * np.cos(phase_acc)
    imag = mag * np.sin(phase_acc)
    return np.concatenate([real, imag], axis=-1).astype(np.float32)
