"""Host-side randomness of the hot path, drawn in the reference's order and passed
explicitly to the GPU (and to the test oracle), so both consume identical draws.

Draw order per output clip: dataset shuffles (pipeline.py:147,154,164)
-> background crop offset (35) -> n_voices (43) -> per voice {gain u (50), offset (69)}
-> n_noises (87) -> per noise {gain u (94), crop offset (103)} -> 6 x {time-mask size
(transforms.py:25), offset (26)} -> {freq-mask size, offset} -> random_merge_aug factors
(data_utils.py:109).
"""
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from .errors import InvalidArgumentError


class ShuffleStream:
    """``Dataset.from_generator(items).repeat().shuffle(len(items))`` as a stream of item
    ids (pipeline.py:143-147): a buffer of ``buffer_size`` ids fed by the endlessly
    repeated sequence 0..n-1; each draw emits a random buffer slot and refills it."""

    def __init__(self, n, rng, buffer_size=None):
        self.n = int(n)
        self.rng = rng
        self.next_up = 0
        size = self.n if buffer_size is None else int(buffer_size)
        self.buf = [self._pull() for _ in range(size)]

    def _pull(self):
        v = self.next_up
        self.next_up = (self.next_up + 1) % self.n
        return v

    def take(self, k):
        out = np.empty(k, np.int32)
        for i in range(k):
            j = int(self.rng.integers(len(self.buf)))
            out[i] = self.buf[j]
            self.buf[j] = self._pull()
        return out


@dataclass
class BatchDraws:
    """All draws of one batch; field meanings as in ``iris_plan`` (include/iris.h)."""
    batch: int
    n_frame: int
    max_voices: int
    max_noises: int
    bg_id: np.ndarray
    bg_offset: np.ndarray
    n_voices: Optional[np.ndarray] = None
    voice_id: Optional[np.ndarray] = None
    voice_u: Optional[np.ndarray] = None       # uniform draw; gain = 10 ** -u in fp32
    voice_gain: Optional[np.ndarray] = None
    voice_offset: Optional[np.ndarray] = None
    n_noises: Optional[np.ndarray] = None
    noise_id: Optional[np.ndarray] = None
    noise_u: Optional[np.ndarray] = None
    noise_gain: Optional[np.ndarray] = None
    noise_offset: Optional[np.ndarray] = None
    time_masks: Optional[np.ndarray] = None    # [B, n, 2] (size, offset)
    freq_masks: Optional[np.ndarray] = None
    merge_factor: Optional[np.ndarray] = None  # [B, n_out-2]
    min_ratio: float = 2 / 3
    min_noise_ratio: float = 1 / 2
    extra: dict = field(default_factory=dict)

    def slice(self, lo, hi):
        """Contiguous sample slice (multi-GPU sharding by sample index)."""
        kw = {}
        for k, v in self.__dict__.items():
            if isinstance(v, np.ndarray):
                kw[k] = v[lo:hi]
            else:
                kw[k] = v
        kw['batch'] = hi - lo
        return BatchDraws(**kw)


def placement(n_frame, padded_len, ratio):
    """pad/len/shift arithmetic of pipeline.py:58-66 (voices) and 95-102 (noises):
    ``pad = T - int32(float32(ratio) * float32(padded_len))``; padded both sides if > 0."""
    pad = int(n_frame) - int(np.int32(np.float32(ratio) * np.float32(padded_len)))
    if pad > 0:
        return pad, int(padded_len) + 2 * pad
    return 0, int(padded_len)


def _randint(rng, high, size=None):
    """Uniform integers in ``[0, high)`` for an ARRAY of exclusive upper bounds: ``floor(u * high)``
    with ``u`` a float64 in [0, 1) (never reaches ``high``; the bias is below 2**-30 for the frame
    and bin counts drawn here).  ``Generator.integers`` with array bounds is ~3x slower, and the
    host draws are what bounds the pipeline once the features stay on the device."""
    high = np.asarray(high)
    u = rng.random(high.shape if size is None else size)
    return (u * high).astype(np.int64)


def draw_batch(rng, batch, n_frame, bg_frames, voice_frames=None, noise_frames=None,
               max_voices=0, max_noises=0, snr=-20, min_ratio=2 / 3, min_noise_ratio=1 / 2,
               n_time_masks=0, time_mask_max=24, n_freq_masks=0, freq_mask_max=16,
               n_bins=257, merge_extra=0, streams=None):
    """Draw one batch (vectorised over clips; draws of different clips are independent, so
    only the per-clip ORDER of the reference matters and it is kept in the layout of the
    arrays).  ``*_frames`` are the per-item frame counts of the registered banks
    (``1 + n_samples // 256``).  ``streams`` = optional dict of ShuffleStream per bank
    (default: ids drawn uniformly, i.e. a shuffle buffer in steady state)."""
    f32 = np.float32
    B, T, V, M = int(batch), int(n_frame), int(max_voices), int(max_noises)
    bg_frames = np.asarray(bg_frames)
    streams = streams or {}

    def ids(name, n, k):
        if name in streams:
            return streams[name].take(k)
        return rng.integers(0, n, size=k, dtype=np.int32)

    d = BatchDraws(batch=B, n_frame=T, max_voices=V, max_noises=M,
                   bg_id=ids('bg', len(bg_frames), B), bg_offset=None,
                   min_ratio=min_ratio, min_noise_ratio=min_noise_ratio)
    bgT = bg_frames[d.bg_id].astype(np.int64)
    tiled = bgT * ((T + bgT - 1) // bgT)
    d.bg_offset = _randint(rng, tiled - T + 1).astype(np.int32)           # random_crop (35)
    if V > 0:
        voice_frames = np.asarray(voice_frames)
        d.voice_id = ids('voice', len(voice_frames), B * V).reshape(B, V)
        d.n_voices = (rng.integers(1, V, size=B, dtype=np.int32) if V > 1
                      else np.ones(B, np.int32))                           # (43)
        vP = voice_frames[d.voice_id].max(axis=1)                          # padded_batch (155)
        pad = T - (f32(min_ratio) * vP.astype(f32)).astype(np.int32)       # (58-59)
        length = np.where(pad > 0, vP + 2 * pad, vP)
        if np.any(length - T <= 0):
            b = int(np.argmax(length - T <= 0))
            raise InvalidArgumentError(
                'clip %d: voice group of padded length %d leaves an empty offset range for '
                'n_frame=%d (pipeline.py:68-69)' % (b, int(vP[b]), T))
        live = np.arange(V)[None, :] < d.n_voices[:, None]
        d.voice_u = (rng.random((B, V), dtype=f32) * f32(-snr / 10)) * live          # (50)
        d.voice_offset = (_randint(rng, (length - T)[:, None], size=(B, V)) * live  # (69)
                          ).astype(np.int32)
        d.voice_gain = np.power(f32(10.), -d.voice_u, dtype=f32)           # pow(10., -u)
    if M > 0:
        noise_frames = np.asarray(noise_frames)
        d.noise_id = ids('noise', len(noise_frames), B * M).reshape(B, M)
        d.n_noises = rng.integers(0, M, size=B, dtype=np.int32)           # (87)
        nP = noise_frames[d.noise_id].max(axis=1)
        pad = T - (f32(min_noise_ratio) * nP.astype(f32)).astype(np.int32)  # (95-96)
        length = np.where(pad > 0, nP + 2 * pad, nP)
        if np.any(length < T):
            raise InvalidArgumentError('noise group shorter than n_frame after padding')
        live = np.arange(M)[None, :] < d.n_noises[:, None]
        d.noise_u = (rng.random((B, M), dtype=f32) * f32(2)) * live        # (94)
        d.noise_offset = (_randint(rng, (length - T + 1)[:, None], size=(B, M)) * live  # (103)
                          ).astype(np.int32)
        d.noise_gain = np.power(f32(10.), -d.noise_u, dtype=f32)
    if n_time_masks:                                                       # transforms.py:25-26
        tm = np.empty((B, n_time_masks, 2), np.int32)
        tm[..., 0] = rng.integers(0, time_mask_max, size=(B, n_time_masks), dtype=np.int32)
        tm[..., 1] = _randint(rng, T - tm[..., 0])
        d.time_masks = tm
    if n_freq_masks:
        fm = np.empty((B, n_freq_masks, 2), np.int32)
        fm[..., 0] = rng.integers(0, freq_mask_max, size=(B, n_freq_masks), dtype=np.int32)
        fm[..., 1] = _randint(rng, n_bins - fm[..., 0])
        d.freq_masks = fm
    if merge_extra:                                                        # data_utils.py:109
        d.merge_factor = f32(0.1) + rng.random((B, merge_extra), dtype=f32) * f32(0.8)
    return d
