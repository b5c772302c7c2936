#!/usr/bin/env python
"""Ramp-up / steady state / tail of the persistent k_fused<FM_MEL> launch, from the per-CTA time
stamps of a -DIRIS_TRACE build of libiris (gpurun_scratch/trace/libiris.so, built by
scripts/build_variants.py).  cfg2 shape, B clips, min-max log-mel.

usage: IRIS_LIB=gpurun_scratch/trace/libiris.so python scripts/trace_fused.py [B]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
TRACE = os.path.abspath('gpurun_out/trace.bin')
os.environ['IRIS_TRACE_FILE'] = TRACE
import numpy as np
import torch
from challenge_b200 import _lib as L
from challenge_b200.engine import Engine
from challenge_b200.plan import draw_batch
from challenge_b200.synth import synthetic_banks

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
GHZ = 1.965
eng = Engine(0)
eng.set_mel(80)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
bgs, voices, labels, noises = synthetic_banks(20202, 2, 64, 256, 64)
bf = eng.register_bank(L.BANK_BG, bgs)
vf = eng.register_bank(L.BANK_VOICE, voices, labels=labels)
nf = eng.register_bank(L.BANK_NOISE, noises)
rng = np.random.default_rng(1)
d = draw_batch(rng, B, 626, bf, vf, nf, max_voices=7, max_noises=2, snr=-20, min_ratio=1,
               n_time_masks=6, n_freq_masks=1)
eng.upload_plan(d)
eng.labels()
o = torch.empty(eng.feature_shape(L.FEAT_LOGMEL_MINMAX), device='cuda')
MODE = L.FEAT_MEL if os.environ.get('TRACE_MODE') == 'mel' else L.FEAT_LOGMEL_MINMAX
for i in range(4):
    flush.fill_(i)
    eng.features(MODE, out=o)
    torch.cuda.synchronize()
tr = np.fromfile(TRACE, dtype=np.uint64).reshape(-1, 64).astype(np.int64)
g_tiles = tr[-1, 0]
tr = tr[:-1]
tr = tr[tr[:, 0] != 0]
n = len(tr)


def q(x):
    x = np.asarray(x, dtype=np.float64)
    return 'min %8.2f  p10 %8.2f  med %8.2f  p90 %8.2f  max %8.2f' % (
        x.min(), np.percentile(x, 10), np.median(x), np.percentile(x, 90), x.max())


def cyc(a, b):
    return (tr[:, a] - tr[:, b]) / GHZ / 1e3   # us


print('B = %d, %d CTAs, IRIS_TAIL1=%s IRIS_TAIL2=%s IRIS_CHUNK=%s' % (B, n, os.environ.get('IRIS_TAIL1'), os.environ.get('IRIS_TAIL2'), os.environ.get('IRIS_CHUNK')))
print('globaltimer, us after the first k_tiles thread')
print('  CTA entry            ', q((tr[:, 0] - g_tiles) / 1e3))
print('  CTA exit (consumer 0)', q((tr[:, 10] - g_tiles) / 1e3))
print('  kernel span (first entry -> last exit): %.2f us' % ((tr[:, 10].max() - tr[:, 0].min()) / 1e3))
print('clock64 of the CTA, us after its entry')
print('  prologue done        ', q(cyc(2, 1)))
print('  grid dependency done ', q(cyc(3, 1)))
print('  first claim returned ', q(cyc(4, 1)))
print('  first bulk copy out  ', q(cyc(5, 1)))
print('  first stage landed   ', q(cyc(6, 1)))
print('  first tile done      ', q(cyc(7, 1)))
print('  producer out of work ', q(cyc(11, 1)))
print('  consumer 0 exit      ', q(cyc(8, 1)))
print('  tiles per CTA        ', q(tr[:, 9]))
ex = cyc(8, 1) + (tr[:, 0] - g_tiles) / 1e3
print('tail: last exit - mean exit = %.2f us, last - median = %.2f us' % (ex.max() - ex.mean(), ex.max() - np.median(ex)))
# steady state: time per tile between stamps 16+5 (10 tiles done) and 16+25 (50 tiles done)
ok = (tr[:, 16 + 25] != 0) & (tr[:, 16 + 5] != 0)
per_tile = (tr[ok, 16 + 25] - tr[ok, 16 + 5]) / 40.0 / GHZ / 1e3
print('steady state us per tile (tiles 10..50)', q(per_tile))
print('  => %d tiles / %d CTAs * median = %.1f us' % (tr[:, 9].sum(), n, tr[:, 9].sum() / n * np.median(per_tile)))
ok = tr[:, 14] != 0
print('tile 21 of a CTA, consumer warp 0 (us): wait+mix', q((tr[ok, 13] - tr[ok, 12]) / GHZ / 1e3))
print('                                        fft     ', q((tr[ok, 14] - tr[ok, 13]) / GHZ / 1e3))
print('                                        epilogue', q((tr[ok, 16 + 11] - tr[ok, 14]) / GHZ / 1e3))
# where the time of a consumer warp goes (sums over all tiles of the CTA)
for w, o in ((0, 48), (5, 54)):
    a = tr[:, o:o + 6].astype(np.float64)
    tiles = tr[:, 9].astype(np.float64)
    tot = a[:, 0] + a[:, 2] + a[:, 3] + a[:, 4]
    print('consumer warp %d, share of its time: wait first stage %.1f%%  wait later stages %.1f%%  mix %.1f%%  fft %.1f%%  '
          'epilogue %.1f%%   (stages per tile %.2f)' % (w, 100 * a[:, 0].sum() / tot.sum(), 100 * a[:, 1].sum() / tot.sum(),
          100 * (a[:, 2] - a[:, 1]).sum() / tot.sum(), 100 * a[:, 3].sum() / tot.sum(), 100 * a[:, 4].sum() / tot.sum(),
          a[:, 5].sum() / tiles.sum()))
    per = tot / tiles / GHZ / 1e3
    fast = per < np.median(per)
    for name, m in (('faster half of the CTAs', fast), ('slower half', ~fast)):
        t = tot[m].sum()
        print('   %-24s %.2f us per tile: wait0 %.1f%% wait+ %.1f%% mix %.1f%% fft %.1f%% epi %.1f%%' % (
            name, per[m].mean(), 100 * a[m, 0].sum() / t, 100 * a[m, 1].sum() / t, 100 * (a[m, 2] - a[m, 1]).sum() / t,
            100 * a[m, 3].sum() / t, 100 * a[m, 4].sum() / t))
if tr[:, 62].any():
    print('post warp of a CTA: items', q(tr[:, 62]), '\n   us waiting for clips', q(tr[:, 60] / GHZ / 1e3),
          '\n   us normalising     ', q(tr[:, 61] / GHZ / 1e3), '\n   us alive           ', q(tr[:, 63] / GHZ / 1e3))
    print('   us per item', q(tr[:, 61] / np.maximum(tr[:, 62], 1) / GHZ / 1e3))
    print('consumer warp 0 after its tiles: items', q(tr[:, 15]), ' us normalising', q(tr[:, 12] / GHZ / 1e3), ' us waiting', q(tr[:, 13] / GHZ / 1e3))
# first tiles: how long do tiles 0-1, 2-3 ... take
for k in (1, 2, 3, 4):
    a = 7 if k == 1 else 16 + k - 1
    print('  tiles %d..%d done at' % (2 * k - 2, 2 * k - 1), q(cyc(16 + k, 1)))
