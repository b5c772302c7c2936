#!/usr/bin/env python
"""Quick timing of the fused mel path (cfg1 / cfg2, C=2, B=256) for the current IRIS_FR.
usage: IRIS_FR=16 python scripts/kbench2.py [modes...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from challenge_b200 import _lib as L
from challenge_b200.engine import Engine
from challenge_b200.plan import draw_batch
from challenge_b200.synth import synthetic_banks

B = 256
want = sys.argv[1:] or ['MEL', 'LOGMEL_MINMAX']
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
C = int(os.environ.get('KB_C', '2'))
bgs, voices, labels, noises = synthetic_banks(20202, C, 64, 256, 64)
bf = eng.register_bank(L.BANK_BG, bgs); vf = eng.register_bank(L.BANK_VOICE, voices, labels=labels)
nf = eng.register_bank(L.BANK_NOISE, noises)
rng = np.random.default_rng(1)
plans = {'cfg1': draw_batch(rng, B, 626, bf),
         'cfg2': draw_batch(rng, B, 626, bf, vf, nf, max_voices=7, max_noises=2, snr=-20, min_ratio=1, n_time_masks=6, n_freq_masks=1)}
modes = dict(MEL=L.FEAT_MEL, LOGMEL_MINMAX=L.FEAT_LOGMEL_MINMAX, COMPLEX=L.FEAT_COMPLEX, MAGPHASE=L.FEAT_MAGPHASE)
for name, d in plans.items():
    eng.upload_plan(d)
    keep = None
    if d.max_voices:
        _, _, keep = eng.labels(); keep = keep.cpu().numpy()
    for m in want:
        out = torch.empty(eng.feature_shape(modes[m]), device='cuda')
        bi, bo = eng.plan_bytes(modes[m], keep)
        med = timeit(lambda: eng.features(modes[m], out=out))
        print('FR=%s C=%d %s %-14s %8.1f us  %6.0f GB/s  frac %.3f' % (os.environ.get('IRIS_FR', 'dflt'), C, name, m, med, (bi + bo) / med / 1e3, (bi + bo) / med / 1e3 / 6454.3), flush=True)
