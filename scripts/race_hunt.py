import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from challenge_b200 import _lib as L
from challenge_b200.engine import Engine
from challenge_b200.plan import draw_batch
from challenge_b200.synth import synthetic_banks
B=256
eng = Engine(0); eng.set_mel(80)
bgs, voices, labels, noises = synthetic_banks(20202, 2, 4, 24, 6)
bf = eng.register_bank(L.BANK_BG, bgs); vf = eng.register_bank(L.BANK_VOICE, voices, labels=labels); nf = eng.register_bank(L.BANK_NOISE, noises)
d = draw_batch(np.random.default_rng(2024), B, 626, bf, vf, nf, max_voices=7, max_noises=2, snr=-20, min_ratio=1, n_time_masks=6, n_freq_masks=1)
eng.upload_plan(d); eng.labels()
for mode in (L.FEAT_MEL, L.FEAT_LOGMEL_MINMAX, L.FEAT_COMPLEX):
    ref = eng.features(mode).clone()
    bad = 0
    for it in range(10):
        x = eng.features(mode)
        ne = (x != ref) & ~(torch.isnan(x) & torch.isnan(ref))
        n = int(ne.sum())
        if n:
            bad += 1
            idx = ne.nonzero()
            print('mode', mode, 'iter', it, 'mismatches', n, 'clips', idx[:, 0].unique().tolist()[:8],
                  'dim1', idx[:, 1].unique().tolist()[:10], 'dim2', idx[:, 2].unique().tolist()[:20], 'dim3', idx[:,3].unique().tolist())
            i0 = tuple(idx[0].tolist()); print('  first', i0, float(x[i0]), float(ref[i0]))
    print('mode', mode, 'bad launches', bad, 'of 10')
