#!/usr/bin/env python
"""Measured float-parity figures behind the bars in tests/test_gpu_parity.py (run on the GPU box):
phase (circular error by magnitude gate, magnitude-weighted error, branch-cut sign mismatches) and
log-magnitude (absolute log error by gate, linear-domain normalised error) for the cfg3 workload."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from conftest import Workload, phase_report, logmag_report
from challenge_b200 import _lib as L
from challenge_b200.engine import Engine
from challenge_b200.plan import draw_batch
from oracle import chain

eng = Engine(0)
eng.set_mel(80)
w = Workload(eng, 4, 20203, n_bg=2, n_voice=16, n_noise=4)
for seed in (33, 34, 35):
    d = draw_batch(np.random.default_rng(seed), 4, 626, w.bg_frames, w.voice_frames, w.noise_frames, max_voices=7,
                   max_noises=2, snr=-20, min_ratio=1, n_time_masks=6, n_freq_masks=1)
    eng.upload_plan(d)
    eng.labels()
    got = eng.features(L.FEAT_MAGPHASE).cpu().numpy()
    ref = chain.dataset_batch(w.o_bg, w.o_voice, w.labels, w.o_noise, d, mode='magphase')[0]
    print('seed', seed, 'phase', phase_report(ref[..., :4], got[..., 4:], ref[..., 4:]))
    got = eng.features(L.FEAT_LOG_MAGPHASE).cpu().numpy()
    ref = chain.dataset_batch(w.o_bg, w.o_voice, w.labels, w.o_noise, d, mode='log_magphase')[0]
    print('seed', seed, 'log-mag', logmag_report(got[..., :4], ref[..., :4]))

# where do the branch-cut flips sit?
d = draw_batch(np.random.default_rng(33), 4, 626, w.bg_frames, w.voice_frames, w.noise_frames, max_voices=7,
               max_noises=2, snr=-20, min_ratio=1, n_time_masks=6, n_freq_masks=1)
eng.upload_plan(d)
eng.labels()
got = eng.features(L.FEAT_MAGPHASE).cpu().numpy()
gc = eng.features(L.FEAT_COMPLEX).cpu().numpy()
ref = chain.dataset_batch(w.o_bg, w.o_voice, w.labels, w.o_noise, d, mode='magphase')[0]
rc = chain.dataset_batch(w.o_bg, w.o_voice, w.labels, w.o_noise, d, mode='complex')[0]
raw = got[..., 4:].astype(np.float64) - ref[..., 4:]
sel = (np.abs(raw) > np.pi) & (ref[..., :4] > 1e-3 * ref[..., :4].max())
idx = np.argwhere(sel)
print('flips', len(idx), 'bins', np.unique(idx[:, 1], return_counts=True))
for b, f, t, c in idx[:12]:
    print(b, f, t, c, 'ref re/im', rc[b, f, t, c], rc[b, f, t, 4 + c], np.signbit(rc[b, f, t, 4 + c]), 'got re/im', gc[b, f, t, c], gc[b, f, t, 4 + c],
          np.signbit(gc[b, f, t, 4 + c]), 'phase ref/got', ref[b, f, t, 4 + c], got[b, f, t, 4 + c])
