"""Oracle restatement of the reference's ``data_utils.py`` (TEST INFRASTRUCTURE ONLY).

numpy fp32 in the reference's op order; the STFT is the reference's own library
call (torchaudio ``Spectrogram(512, power=None)``, data_utils.py:17) with a
``torch.stft`` spelling of the same parameters as the fallback.
"""
import numpy as np

from . import EPSILON
from .transforms import mask

N_FFT = 512
HOP = 256


def normalize(wav):
    """data_utils.py:32-34 -- ``wav / (10 * sqrt(mean(wav**2)))`` over ALL channels."""
    import torch
    wav = torch.as_tensor(np.asarray(wav, dtype=np.float32))
    rms = torch.sqrt(torch.mean(torch.pow(wav, 2))) * 10
    return (wav / rms).numpy()


def stft(wav):
    """data_utils.py:17,23 -- complex STFT ``[C, N] -> complex64 [C, 257, T]``.

    n_fft 512, win 512 periodic Hann, hop 256, center=True reflect pad,
    onesided, un-normalised.
    """
    import torch
    wav = torch.as_tensor(np.ascontiguousarray(wav, dtype=np.float32))
    try:
        import torchaudio
        spec = torchaudio.transforms.Spectrogram(N_FFT, power=None)(wav)
    except Exception:  # pragma: no cover - torchaudio is present in this image
        spec = torch.stft(wav, N_FFT, hop_length=HOP, win_length=N_FFT,
                          window=torch.hann_window(N_FFT), center=True,
                          pad_mode='reflect', normalized=False, onesided=True,
                          return_complex=True)
    return spec.numpy()


def stft_f64(wav):
    """Independent float64 numpy restatement of SURVEY A.1 (cross-check only)."""
    wav = np.asarray(wav, dtype=np.float64)
    C, N = wav.shape
    T = 1 + N // HOP
    pad = np.pad(wav, ((0, 0), (N_FFT // 2, N_FFT // 2)), mode='reflect')
    n = np.arange(N_FFT)
    w = 0.5 - 0.5 * np.cos(2 * np.pi * n / N_FFT)
    idx = HOP * np.arange(T)[:, None] + n[None, :]
    frames = pad[:, idx] * w  # [C, T, 512]
    return np.fft.rfft(frames, axis=-1).transpose(0, 2, 1)  # [C, 257, T]


def spec_layout(spec_c):
    """data_utils.py:25-27 -- complex ``[C,F,T]`` -> real ``[F,T,2C]``.

    Old torchaudio returned a real view ``[C,F,T,2]``; the reference transposes
    (1,2,3,0) -> ``[F,T,2,C]`` and flattens the last two dims, so
    ``[..., :C]`` = real and ``[..., C:]`` = imag.
    """
    ri = np.stack([spec_c.real, spec_c.imag], axis=-1).astype(np.float32)  # [C,F,T,2]
    out = ri.transpose(1, 2, 3, 0)
    return np.ascontiguousarray(out.reshape((*out.shape[:2], -1)))


def load_wav_array(wav, do_normalize=True):
    """data_utils.py:9-29 minus file decode / resample (16 kHz input assumed)."""
    wav = np.asarray(wav, dtype=np.float32)
    if do_normalize:
        wav = normalize(wav)
    return spec_layout(stft(wav))


def safe_div(x, y, eps=EPSILON):
    """utils.py:114-116."""
    return x / np.maximum(y, np.float32(eps))


def minmax(x, y=None):
    """data_utils.py:37-47 -- per-sample (axis 0 = batch) global min-max."""
    x = np.asarray(x, dtype=np.float32)
    axis = tuple(range(1, x.ndim))
    x_max = x.max(axis=axis, keepdims=True)
    x_min = x.min(axis=axis, keepdims=True)
    x = safe_div(x - x_min, x_max - x_min)
    if y is not None:
        return x, y
    return x


def log_on_mel(mel, labels=None):
    """data_utils.py:50-55."""
    mel = np.log(np.asarray(mel, dtype=np.float32) + np.float32(EPSILON))
    if labels is not None:
        return mel, labels
    return mel


def augment(specs, labels, time_axis=-2, freq_axis=-3, *, time_draws, freq_draws):
    """data_utils.py:58-61 -- 6 time masks (<24) then 1 freq mask (<16)."""
    specs = mask(specs, axis=time_axis, max_mask_size=24, n_mask=6, draws=time_draws)
    specs = mask(specs, axis=freq_axis, max_mask_size=16, draws=freq_draws)
    return specs, labels


def to_frame_labels(x, y):
    """data_utils.py:64-70."""
    return x, np.sum(y, axis=-3, dtype=np.float32)


def mono_chan(x, y=None):
    """data_utils.py:73-76 -- broadcast quirk kept: ``x[..., :1] + x[..., 1:]``."""
    if y is not None:
        return x[..., :1] + x[..., 1:], y
    return x


def stereo_mono(x, y=None):
    """data_utils.py:79-82."""
    out = np.concatenate([x[..., :2], x[..., :1] + x[..., 1:2],
                          x[..., 2:4], x[..., 2:3] + x[..., 3:4]], -1)
    if y is None:
        return out
    return out, y


def _avg_pool1d_same(y, r):
    """Keras AveragePooling1D(r, r, 'same') [TF-sem]: average over VALID cells."""
    B, T, K = y.shape
    out_len = -(-T // r)
    total_pad = max((out_len - 1) * r + r - T, 0)
    pad_left = total_pad // 2
    out = np.zeros((B, out_len, K), np.float32)
    for i in range(out_len):
        lo = max(i * r - pad_left, 0)
        hi = min(i * r - pad_left + r, T)
        out[:, i] = y[:, lo:hi].sum(axis=1, dtype=np.float32) / np.float32(hi - lo)
    return out


def label_downsample(resolution=32):
    """data_utils.py:85-97 -- note ``[:resolution]`` slices the BATCH axis (quirk)."""
    def _label_downsample(x, y):
        if isinstance(y, (list, tuple)):
            y_ = _avg_pool1d_same(np.asarray(y[0], np.float32), resolution)
            y_ = (y_ >= 0.5).astype(np.float32)[:resolution]
            y = (y_,) + tuple(y[1:])
        else:
            y = _avg_pool1d_same(np.asarray(y, np.float32), resolution)
            y = (y >= 0.5).astype(np.float32)[:resolution]
        return x, y
    return _label_downsample


def random_merge_aug(number):
    """data_utils.py:100-117 -- ``factor`` drawn U(0.1, 0.9), shape (1,1,number-2)."""
    def _random_merge_aug(x, y=None, *, factor):
        chan = x.shape[-1] // 2
        if chan != 2:
            raise ValueError('This augment can be used in 2 channel audio')
        real = x[..., :chan]
        imag = x[..., chan:]
        factor = np.asarray(factor, np.float32).reshape((1,) * (x.ndim - 1) + (number - chan,))
        aug_real = factor * np.repeat(real[..., :1], number - chan, -1) \
            + np.sqrt(np.float32(1) - factor) * np.repeat(real[..., 1:], number - chan, -1)
        real = np.concatenate([real, aug_real], -1)
        imag = np.concatenate(
            [imag, np.repeat(imag[..., :1] + imag[..., 1:], number - chan, -1)], -1)
        out = np.concatenate([real, imag], -1).astype(np.float32)
        if y is not None:
            return out, y
        return out
    return _random_merge_aug


def multiply_label(multiply_factor):
    """data_utils.py:120-123."""
    def _multiply_label(x, y):
        return x, y * np.float32(multiply_factor)
    return _multiply_label


def stft_filter(filter_num):
    """data_utils.py:126-136 -- zero bins 1..filter_num by MULTIPLICATION."""
    def _stft_filter(x, y=None):
        m = np.ones((x.shape[0],) + (1,) * (x.ndim - 1), x.dtype)
        m[1:1 + filter_num] = 0
        x = x * m
        if y is None:
            return x
        return x, y
    return _stft_filter


def speech_enhancement_preprocess(x, y=None):
    """data_utils.py:139-148."""
    x = x[1:, ..., :x.shape[-1] // 2]
    if y is None:
        return x
    y = (np.sum(y[0], axis=-3), y[1][1:, ..., :x.shape[-1] // 2],
         y[2][1:, ..., :x.shape[-1] // 2])
    return x, y


# ---- torchaudio.compliance.kaldi.resample_waveform (data_utils.py:20-21) ----
# THIRD-PARTY ALGORITHM, not under /root/reference: torchaudio (requirements.txt:5, unpinned; the
# reference's API usage -- Spectrogram(power=None) returning a real [..., 2] view -- implies
# <= 0.8, where resample_waveform is the port of Kaldi's LinearResample restated below; since 0.9
# the same filter lives in torchaudio.functional.resample, which tests/test_oracle_golden.py runs
# here as the pin: identical window, sinc, scaling, zero padding and output length).
def _lr_num_output_samples(input_num_samp, samp_rate_in, samp_rate_out):
    """kaldi.py::_get_num_LR_output_samples: outputs whose time lies in [0, len / orig_freq)."""
    import math
    samp_in, samp_out = int(samp_rate_in), int(samp_rate_out)
    tick_freq = samp_in * samp_out // math.gcd(samp_in, samp_out)
    ticks_per_input_period = tick_freq // samp_in
    interval_length_in_ticks = input_num_samp * ticks_per_input_period
    if interval_length_in_ticks <= 0:
        return 0
    ticks_per_output_period = tick_freq // samp_out
    last_output_samp = interval_length_in_ticks // ticks_per_output_period
    if last_output_samp * ticks_per_output_period == interval_length_in_ticks:
        last_output_samp -= 1
    return last_output_samp + 1


def _lr_indices_and_weights(orig_freq, new_freq, output_samples_in_unit, window_width,
                            lowpass_cutoff, lowpass_filter_width):
    """kaldi.py::_get_LR_indices_and_weights in float32, op for op: for every output phase the
    first input sample of its window and the windowed-sinc weights."""
    f32 = np.float32
    output_t = np.arange(output_samples_in_unit, dtype=f32) / f32(new_freq)
    min_t = output_t - f32(window_width)
    max_t = output_t + f32(window_width)
    min_input_index = np.ceil(min_t * f32(orig_freq))
    max_input_index = np.floor(max_t * f32(orig_freq))
    num_indices = max_input_index - min_input_index + 1
    max_weight_width = int(num_indices.max())
    j = np.arange(max_weight_width, dtype=f32)[None, :]
    input_index = min_input_index[:, None] + j
    delta_t = (input_index / f32(orig_freq)) - output_t[:, None]
    weights = np.zeros_like(delta_t)
    inside = np.abs(delta_t) < f32(window_width)
    # raised-cosine (Hanning) window of width window_width
    weights[inside] = f32(0.5) * (1 + np.cos(
        f32(2 * np.pi * lowpass_cutoff / lowpass_filter_width) * delta_t[inside], dtype=f32))
    zero = delta_t == 0
    nz = ~zero
    weights[nz] *= np.sin(f32(2 * np.pi * lowpass_cutoff) * delta_t[nz], dtype=f32) / (
        f32(np.pi) * delta_t[nz])
    weights[zero] *= f32(2 * lowpass_cutoff)
    weights /= f32(orig_freq)
    return min_input_index.astype(np.int64), weights.astype(f32)


def resample_waveform(wav, orig_freq, new_freq, lowpass_filter_width=6):
    """kaldi.resample_waveform: out[c, u*U_out + i] = sum_j w[i, j] * wav[c, first[i] + u*U_in + j],
    zero outside the clip (the port's conv1d with stride U_in per output phase i, zero-padded)."""
    import math
    wav = np.asarray(wav, np.float32)
    assert wav.ndim == 2 and orig_freq > 0 and new_freq > 0
    if int(orig_freq) == int(new_freq):
        return wav.copy()
    min_freq = min(orig_freq, new_freq)
    lowpass_cutoff = 0.99 * 0.5 * min_freq
    assert lowpass_cutoff * 2 <= min_freq
    base_freq = math.gcd(int(orig_freq), int(new_freq))
    u_in = int(orig_freq) // base_freq
    u_out = int(new_freq) // base_freq
    window_width = lowpass_filter_width / (2.0 * lowpass_cutoff)
    first, weights = _lr_indices_and_weights(orig_freq, new_freq, u_out, window_width,
                                             lowpass_cutoff, lowpass_filter_width)
    C, n_in = wav.shape
    n_out = _lr_num_output_samples(n_in, orig_freq, new_freq)
    W = weights.shape[1]
    n_units = (n_out + u_out - 1) // u_out
    lo = int(min(first.min(), 0))
    hi = int(first.max()) + (n_units - 1) * u_in + W
    padded = np.zeros((C, max(hi, n_in) - lo), np.float32)
    padded[:, -lo:-lo + n_in] = wav
    out = np.zeros((C, n_units * u_out), np.float32)
    units = np.arange(n_units) * u_in
    for i in range(u_out):
        acc = np.zeros((C, n_units), np.float32)
        base = units + int(first[i]) - lo
        for jj in range(W):
            acc += weights[i, jj] * padded[:, base + jj]
        out[:, i::u_out] = acc
    return out[:, :n_out]
