#!/usr/bin/env python
"""A small pass over every kernel of the step path for compute-sanitizer (memcheck / racecheck /
synccheck): 3 clips x 120 frames, 2 and 4 channels, labels (+ tile lists) -> features in every
mode -> metric counts, plus the one-call step.  usage:
  compute-sanitizer --tool racecheck python scripts/sanitize_small.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from challenge_b200 import _lib as L
from challenge_b200.engine import Engine
from challenge_b200.plan import draw_batch
from challenge_b200.synth import synthetic_banks

for C in (2, 4):
    eng = Engine(0)
    eng.set_mel(80)
    bgs, voices, labels, noises = synthetic_banks(5 + C, C, n_bg=2, n_voice=6, n_noise=2, bg_seconds=2.0)
    bf = eng.register_bank(L.BANK_BG, bgs)
    vf = eng.register_bank(L.BANK_VOICE, voices, labels=labels)
    nf = eng.register_bank(L.BANK_NOISE, noises)
    d = draw_batch(np.random.default_rng(C), 3, 120, bf, vf, nf, max_voices=4, max_noises=2, snr=-20, min_ratio=2 / 3,
                   n_time_masks=6, n_freq_masks=1)
    eng.upload_plan(d)
    for mode in (L.FEAT_LOGMEL_MINMAX, L.FEAT_MEL, L.FEAT_COMPLEX, L.FEAT_MAGPHASE, L.FEAT_LOG_MAGPHASE):
        frame, _, keep = eng.labels()
        x = eng.features(mode)
        x = eng.features(mode)
    yp = torch.clamp(frame + 0.3 * torch.randn_like(frame), 0, 1)
    eng.metric_counts(frame, yp)
    torch.cuda.synchronize()
    print('C = %d ok' % C, float(x.abs().max()))
    eng.close()
print('done')
