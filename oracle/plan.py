"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the host planner (libiris'
``iris_draw_batch``): one block of uniforms per batch -> the draws of merge_complex_specs /
mask / random_merge_aug in the reference's per-clip order.

Layout of a clip's uniforms: bg id, bg crop offset (pipeline.py:35), V voice ids, n_voices
(43), V x {gain u (50), offset (69)}, M noise ids, n_noises (87), M x {gain u (94), crop
offset (103)}, n_time_masks x {size (transforms.py:25), offset (26)}, n_freq_masks x {size,
offset}, merge_extra factors (data_utils.py:109).  Integers are ``floor(u * range)``.
"""
import numpy as np


class ShuffleStreamPy:
    """``Dataset.from_generator(items).repeat().shuffle(buffer)`` (pipeline.py:143-147)."""

    def __init__(self, n, buffer_size=None):
        self.n = int(n)
        self.next_up = 0
        self.buf = [self._pull() for _ in range(self.n if not buffer_size else int(buffer_size))]

    def _pull(self):
        v = self.next_up
        self.next_up = (self.next_up + 1) % self.n
        return v

    def take(self, u):
        j = min(int(u * len(self.buf)), len(self.buf) - 1)
        v = self.buf[j]
        self.buf[j] = self._pull()
        return v


def uniforms_per_clip(V, M, n_tm, n_fm, merge_extra):
    return 2 + (3 * V + 1 if V else 0) + (3 * M + 1 if M else 0) + 2 * n_tm + 2 * n_fm + merge_extra


def _below(u, rng_):
    rng_ = np.asarray(rng_, np.int64)
    v = np.floor(np.asarray(u, np.float64) * rng_).astype(np.int64)
    return np.where(rng_ <= 1, 0, np.minimum(v, rng_ - 1))


def _unit32(u):
    f = np.asarray(u, np.float64).astype(np.float32)
    return np.minimum(f, np.float32(1) - np.float32(2) ** -24)


def draws_from_uniforms(u, n_frame, bg_frames, voice_frames=None, noise_frames=None, max_voices=0,
                        max_noises=0, snr=-20, min_ratio=2 / 3, min_noise_ratio=1 / 2, n_time_masks=0,
                        time_mask_max=24, n_freq_masks=0, freq_mask_max=16, n_bins=257, merge_extra=0,
                        streams=None):
    """-> dict of arrays named like ``iris_draws`` (include/iris.h)."""
    f32 = np.float32
    u = np.asarray(u, np.float64)
    B, T, V, M = u.shape[0], int(n_frame), int(max_voices), int(max_noises)
    assert u.shape[1] == uniforms_per_clip(V, M, n_time_masks, n_freq_masks, merge_extra)
    streams = streams or {}
    bg_frames = np.asarray(bg_frames, np.int64)
    out = {}
    o = 0

    def ids(name, col, n):
        if name in streams:
            return np.array([streams[name].take(x) for x in col], np.int32)
        return _below(col, n).astype(np.int32)

    # clip-major consumption of the streams: bg, then the V voices, then the M noises of a clip;
    # the three streams are independent, so only the order inside each bank matters
    out['bg_id'] = ids('bg', u[:, o], len(bg_frames)); o += 1
    bgT = bg_frames[out['bg_id']]
    tiled = bgT * ((T + bgT - 1) // bgT)
    out['bg_offset'] = _below(u[:, o], tiled - T + 1).astype(np.int32); o += 1
    if V:
        vf = np.asarray(voice_frames, np.int64)
        out['voice_id'] = ids('voice', u[:, o:o + V].reshape(-1), len(vf)).reshape(B, V); o += V
        out['n_voices'] = ((1 + _below(u[:, o], V - 1)) if V > 1 else np.ones(B)).astype(np.int32); o += 1
        vP = vf[out['voice_id']].max(axis=1)
        pad = T - (f32(min_ratio) * vP.astype(f32)).astype(np.int32)
        length = np.where(pad > 0, vP + 2 * pad, vP)
        if np.any(length - T <= 0):
            raise ValueError('empty voice offset range (pipeline.py:68-69)')
        live = np.arange(V)[None, :] < out['n_voices'][:, None]
        pairs = u[:, o:o + 2 * V].reshape(B, V, 2); o += 2 * V
        out['voice_u'] = np.where(live, _unit32(pairs[..., 0]) * f32(-float(f32(snr)) / 10.0), f32(0)).astype(f32)
        out['voice_offset'] = np.where(live, _below(pairs[..., 1], (length - T)[:, None]), 0).astype(np.int32)
        out['voice_gain'] = np.power(f32(10.), -out['voice_u'], dtype=f32)
    if M:
        nf = np.asarray(noise_frames, np.int64)
        out['noise_id'] = ids('noise', u[:, o:o + M].reshape(-1), len(nf)).reshape(B, M); o += M
        out['n_noises'] = _below(u[:, o], M).astype(np.int32); o += 1
        nP = nf[out['noise_id']].max(axis=1)
        pad = T - (f32(min_noise_ratio) * nP.astype(f32)).astype(np.int32)
        length = np.where(pad > 0, nP + 2 * pad, nP)
        if np.any(length < T):
            raise ValueError('noise group shorter than n_frame after padding')
        live = np.arange(M)[None, :] < out['n_noises'][:, None]
        pairs = u[:, o:o + 2 * M].reshape(B, M, 2); o += 2 * M
        out['noise_u'] = np.where(live, _unit32(pairs[..., 0]) * f32(2), f32(0)).astype(f32)
        out['noise_offset'] = np.where(live, _below(pairs[..., 1], (length - T + 1)[:, None]), 0).astype(np.int32)
        out['noise_gain'] = np.power(f32(10.), -out['noise_u'], dtype=f32)
    for name, n, mx, total in (('time_masks', n_time_masks, time_mask_max, T),
                               ('freq_masks', n_freq_masks, freq_mask_max, n_bins)):
        if n:
            pairs = u[:, o:o + 2 * n].reshape(B, n, 2); o += 2 * n
            m = np.empty((B, n, 2), np.int32)
            m[..., 0] = _below(pairs[..., 0], mx)
            m[..., 1] = _below(pairs[..., 1], total - m[..., 0])
            out[name] = m
    if merge_extra:
        out['merge_factor'] = (f32(0.1) + _unit32(u[:, o:o + merge_extra]) * f32(0.8)).astype(f32)
        o += merge_extra
    assert o == u.shape[1]
    return out
