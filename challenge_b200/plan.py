"""Host-side randomness of the hot path, drawn in the reference's order and passed
explicitly to the GPU (and to the test oracle), so both consume identical draws.

The host supplies ONE block of uniforms per batch (``rng.random((B, n_u))``); libiris'
host planner (``iris_draw_batch``, csrc/iris_step.cu) turns it into the draws of an
``iris_plan`` with the reference's placement arithmetic.  Per clip the uniforms are consumed
in the reference's draw order: dataset shuffles (pipeline.py:147,154,164)
-> background crop offset (35) -> n_voices (43) -> per voice {gain u (50), offset (69)}
-> n_noises (87) -> per noise {gain u (94), crop offset (103)} -> 6 x {time-mask size
(transforms.py:25), offset (26)} -> {freq-mask size, offset} -> random_merge_aug factors
(data_utils.py:109).  ``iris_step`` runs the same planner inside the one-call batch step.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _lib as L
from .errors import InvalidArgumentError


class ShuffleStream:
    """``Dataset.from_generator(items).repeat().shuffle(len(items))`` as a stream of item
    ids (pipeline.py:143-147): a buffer of ``buffer_size`` ids fed by the endlessly
    repeated sequence 0..n-1; each draw emits a random buffer slot and refills it.  The
    state lives in libiris (``iris_shuffle``) so that ``iris_step`` advances it from C."""

    def __init__(self, n, rng, buffer_size=None):
        self.n = int(n)
        self.rng = rng
        self._h = C.c_void_p()
        L.check(L.load().iris_shuffle_create(self.n, int(buffer_size or 0), C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def take(self, k, uniforms=None):
        u = np.ascontiguousarray(self.rng.random(int(k)) if uniforms is None else uniforms, np.float64)
        out = np.empty(len(u), np.int32)
        L.check(L.load().iris_shuffle_take(self._h, u.ctypes.data, len(u), out.ctypes.data))
        return out

    def __del__(self):
        try:
            if self._h:
                L.load().iris_shuffle_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


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
            elif k == 'extra':
                kw[k] = {n: (a[lo:hi] if isinstance(a, np.ndarray) else a) for n, a in v.items()}
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


def draw_config(batch, n_frame, max_voices=0, max_noises=0, snr=-20, min_ratio=2 / 3,
                min_noise_ratio=1 / 2, n_time_masks=0, time_mask_max=24, n_freq_masks=0,
                freq_mask_max=16, n_bins=257, merge_extra=0):
    """``iris_draw_config`` (include/iris.h) of a batch."""
    return L.IrisDrawConfig(int(batch), int(n_frame), int(max_voices), int(max_noises), float(min_ratio),
                            float(min_noise_ratio), float(snr), int(n_time_masks), int(time_mask_max),
                            int(n_freq_masks), int(freq_mask_max), int(n_bins), int(merge_extra))


def uniforms_per_clip(cfg):
    return int(L.load().iris_draw_uniforms_per_clip(C.byref(cfg)))


def stream_handles(streams):
    """ctypes ``iris_shuffle*[3]`` (bg, voice, noise) from a dict of ShuffleStream, or None."""
    if not streams:
        return None
    arr = (C.c_void_p * 3)()
    for i, k in enumerate(('bg', 'voice', 'noise')):
        arr[i] = streams[k].handle if k in streams else None
    return arr


def draws_from_uniforms(cfg, uniforms, bg_frames, voice_frames=None, noise_frames=None, streams=None):
    """The host planner of libiris on a block of uniforms ``[B, uniforms_per_clip(cfg)]``."""
    B, T, V, M = cfg.batch, cfg.n_frame, cfg.max_voices, cfg.max_noises
    u = np.ascontiguousarray(uniforms, np.float64)
    assert u.shape == (B, uniforms_per_clip(cfg)), (u.shape, B, uniforms_per_clip(cfg))
    i32, f32 = np.int32, np.float32
    d = BatchDraws(batch=B, n_frame=T, max_voices=V, max_noises=M, bg_id=np.zeros(B, i32),
                   bg_offset=np.zeros(B, i32), min_ratio=cfg.min_ratio, min_noise_ratio=cfg.min_noise_ratio)
    if V > 0:
        d.n_voices, d.voice_id = np.zeros(B, i32), np.zeros((B, V), i32)
        d.voice_u, d.voice_gain, d.voice_offset = np.zeros((B, V), f32), np.zeros((B, V), f32), np.zeros((B, V), i32)
    if M > 0:
        d.n_noises, d.noise_id = np.zeros(B, i32), np.zeros((B, M), i32)
        d.noise_u, d.noise_gain, d.noise_offset = np.zeros((B, M), f32), np.zeros((B, M), f32), np.zeros((B, M), i32)
    if cfg.n_time_masks:
        d.time_masks = np.zeros((B, cfg.n_time_masks, 2), i32)
    if cfg.n_freq_masks:
        d.freq_masks = np.zeros((B, cfg.n_freq_masks, 2), i32)
    if cfg.merge_extra:
        d.merge_factor = np.zeros((B, cfg.merge_extra), f32)
    out = L.IrisDraws()
    for name, _ in L.IrisDraws._fields_:
        a = getattr(d, name)
        if a is not None:
            setattr(out, name, a.ctypes.data_as(L._f32p if a.dtype == np.float32 else L._i32p))

    def frames(a):
        return None if a is None else np.ascontiguousarray(a, np.int32)
    bf, vf, nf = frames(bg_frames), frames(voice_frames), frames(noise_frames)
    L.check(L.load().iris_draw_batch(
        C.byref(cfg), bf.ctypes.data, len(bf), vf.ctypes.data if vf is not None else None,
        len(vf) if vf is not None else 0, nf.ctypes.data if nf is not None else None,
        len(nf) if nf is not None else 0, stream_handles(streams), u.ctypes.data, C.byref(out)))
    d.extra['uniforms'] = u
    return d


def draw_batch(rng, batch, n_frame, bg_frames, voice_frames=None, noise_frames=None,
               max_voices=0, max_noises=0, snr=-20, min_ratio=2 / 3, min_noise_ratio=1 / 2,
               n_time_masks=0, time_mask_max=24, n_freq_masks=0, freq_mask_max=16,
               n_bins=257, merge_extra=0, streams=None):
    """Draw one batch: one block of uniforms from ``rng`` through the host planner.
    ``*_frames`` are the per-item frame counts of the registered banks
    (``1 + n_samples // 256``).  ``streams`` = optional dict of ShuffleStream per bank
    (default: ids drawn uniformly, i.e. a shuffle buffer in steady state)."""
    if max_voices and voice_frames is None:
        max_voices = 0
    if max_noises and noise_frames is None:
        max_noises = 0
    cfg = draw_config(batch, n_frame, max_voices, max_noises, snr, min_ratio, min_noise_ratio,
                      n_time_masks, time_mask_max, n_freq_masks, freq_mask_max, n_bins, merge_extra)
    u = rng.random((int(batch), uniforms_per_clip(cfg)))
    return draws_from_uniforms(cfg, u, bg_frames, voice_frames, noise_frames, streams)
