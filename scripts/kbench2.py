import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from challenge_b200 import _lib as L
from challenge_b200.engine import Engine
from challenge_b200.plan import draw_batch
from challenge_b200.synth import synthetic_banks
B=256
eng = Engine(0); eng.set_mel(80)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
bgs, voices, labels, noises = synthetic_banks(20202, 2, 64, 256, 64)
bf = eng.register_bank(L.BANK_BG, bgs); vf = eng.register_bank(L.BANK_VOICE, voices, labels=labels); nf = eng.register_bank(L.BANK_NOISE, noises)
d = draw_batch(np.random.default_rng(1), B, 626, bf, vf, nf, max_voices=7, max_noises=2, snr=-20, min_ratio=1, n_time_masks=6, n_freq_masks=1)
eng.upload_plan(d); eng.labels()
mode = int(os.environ.get('MODE', L.FEAT_LOGMEL_MINMAX))
out = torch.empty(eng.feature_shape(mode), device='cuda')
ts=[]
for i in range(13):
    flush.fill_(i)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.features(mode, out=out); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b)*1e3)
print('IRIS_DEBUG=%s mode %d median %.1f us min %.1f' % (os.environ.get('IRIS_DEBUG','0'), mode, np.median(ts[3:]), min(ts[3:])))

