import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from challenge_b200.engine import Engine
from challenge_b200 import _lib as L
from conftest import Workload
from challenge_b200.plan import draw_batch
from oracle import chain
eng = Engine(0); eng.set_mel(80)
w = Workload(eng, 2, 20202)
T = 626
rng = np.random.default_rng(T)
d = draw_batch(rng, 8, T, w.bg_frames, w.voice_frames, w.noise_frames, max_voices=7, max_noises=2, snr=-20, min_ratio=1, n_time_masks=6, n_freq_masks=1)
eng.upload_plan(d)
got = eng.features(L.FEAT_COMPLEX).cpu().numpy()
ref = chain.dataset_batch(w.o_bg, w.o_voice, w.labels, w.o_noise, d, mode='complex')[0]
a = got == 0; b = ref == 0
mm = np.argwhere(a != b)
print('mismatches', len(mm), 'got zeros', a.sum(), 'ref zeros', b.sum())
for idx in mm[:30]:
    print(idx, got[tuple(idx)], ref[tuple(idx)])
print('time masks', d.time_masks[mm[0][0]] if len(mm) else None, 'freq', d.freq_masks[mm[0][0]] if len(mm) else None)
