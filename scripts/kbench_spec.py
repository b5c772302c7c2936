#!/usr/bin/env python
"""Spectrogram-bank path (k_specmix) at cfg2 shape: time per mode, algorithmic GB/s vs the HBM peak."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from challenge_b200 import _lib as L
from challenge_b200.engine import Engine
from challenge_b200.plan import draw_batch
from challenge_b200.synth import synthetic_banks
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'] if os.path.exists('MEASURED_PEAKS.json') else 6446.3
eng = Engine(0); eng.set_mel(80)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def timeit(fn, n=10):
    for _ in range(3): fn()
    ts = []
    for i in range(n):
        flush.fill_(i & 255)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))
bgs, voices, labels, noises = synthetic_banks(20202, 2, 64, 256, 64)
spec = lambda ws: [eng.stft(w, normalize=True).cpu().numpy() for w in ws]
bf = eng.register_bank(L.BANK_BG, spec(bgs)); vf = eng.register_bank(L.BANK_VOICE, spec(voices), labels=labels)
nf = eng.register_bank(L.BANK_NOISE, spec(noises))
d = draw_batch(np.random.default_rng(1), B, 626, bf, vf, nf, max_voices=7, max_noises=2, snr=-20, min_ratio=1,
               n_time_masks=6, n_freq_masks=1)
eng.upload_plan(d)
_, _, keep = eng.labels(); keep = keep.cpu().numpy()
for name, mode in (('COMPLEX', L.FEAT_COMPLEX), ('MAGPHASE', L.FEAT_MAGPHASE), ('LOGMEL_MINMAX', L.FEAT_LOGMEL_MINMAX)):
    out = torch.empty(eng.feature_shape(mode), device='cuda')
    bi, bo = eng.plan_bytes(mode, keep)
    us = timeit(lambda: eng.features(mode, out=out))
    print('spec banks C=2 B=%d cfg2 %-14s %8.1f us  %7.1f kclips/s  alg %7.1f MB  %6.0f GB/s (%.3f of %.0f)' % (
        B, name, us, B / us * 1e3, (bi + bo) / 1e6, (bi + bo) / us / 1e3, (bi + bo) / us / 1e3 / peak, peak))
