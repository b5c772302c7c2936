#!/usr/bin/env python
"""Kernel micro-benchmark: times iris_features() (k_tiles + k_fused) per mode with CUDA events.
usage: python scripts/kbench.py [B] [iters] [C: only this channel count]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from challenge_b200 import _lib as L
from challenge_b200.engine import Engine
from challenge_b200.plan import draw_batch
from challenge_b200.synth import synthetic_banks

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
eng = Engine(0); eng.set_mel(80)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')

def timeit(fn, n=iters):
    for _ in range(3): fn()
    ts = []
    for i in range(n):
        flush.fill_(i & 255)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(np.min(ts))

only_c = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for C in ((only_c,) if only_c else (2, 4)):
    bgs, voices, labels, noises = synthetic_banks(20202, C, 64, 256, 64)
    bf = eng.register_bank(L.BANK_BG, bgs); vf = eng.register_bank(L.BANK_VOICE, voices, labels=labels)
    nf = eng.register_bank(L.BANK_NOISE, noises)
    rng = np.random.default_rng(1)
    plans = {
        'cfg1 (bg only, no masks)': draw_batch(rng, B, 626, bf),
        'cfg2 (mix + masks)': draw_batch(rng, B, 626, bf, vf, nf, max_voices=7, max_noises=2, snr=-20, min_ratio=1, n_time_masks=6, n_freq_masks=1),
    }
    for name, d in plans.items():
        eng.upload_plan(d)
        keep = None
        if d.max_voices:
            _, _, keep = eng.labels(); keep = keep.cpu().numpy()
        modes = [('MEL', L.FEAT_MEL), ('LOGMEL', L.FEAT_LOGMEL), ('LOGMEL_MINMAX', L.FEAT_LOGMEL_MINMAX)]
        if C == 4: modes += [('MAGPHASE', L.FEAT_MAGPHASE), ('COMPLEX', L.FEAT_COMPLEX)]
        if C == 2: modes += [('COMPLEX', L.FEAT_COMPLEX)]
        for mname, mode in modes:
            out = torch.empty(eng.feature_shape(mode), device='cuda')
            bi, bo = eng.plan_bytes(mode, keep)
            med, mn = timeit(lambda: eng.features(mode, out=out))
            print('C=%d B=%d %-26s %-14s median %8.1f us  min %8.1f us  %7.1f kclips/s  alg %6.1f MB  %6.0f GB/s (%.3f of 6650)' % (
                C, B, name, mname, med, mn, B / med * 1e3, (bi + bo) / 1e6, (bi + bo) / med / 1e3, (bi + bo) / med / 1e3 / 6650))
    if B > 256: break
