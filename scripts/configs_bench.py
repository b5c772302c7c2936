#!/usr/bin/env python
"""BASELINE.json configs at their stated sizes on ONE B200 (device-timed with CUDA events, L2 flushed
between iterations, median of n): labels + features (+ metric counts), the algorithmic bytes of
SURVEY.md 8d and the fraction of the measured HBM peak, plus the input-bound check of configs[4]
with a SUBSTITUTE consumer (TensorFlow / sj_train.py cannot run here).
usage: python scripts/configs_bench.py [iters]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from challenge_b200 import _lib as L
from challenge_b200.engine import Engine
from challenge_b200.plan import draw_batch
from challenge_b200.synth import synthetic_banks

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 7
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    PEAK = 6650.0
eng = Engine(0); eng.set_mel(80)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')

def timeit(fn, n=iters):
    for _ in range(2): fn()
    ts = []
    for i in range(n):
        flush.fill_(i & 255)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))

def banks(C, seed):
    bgs, voices, labels, noises = synthetic_banks(seed, C, 64, 256, 64)
    bf = eng.register_bank(L.BANK_BG, bgs); vf = eng.register_bank(L.BANK_VOICE, voices, labels=labels)
    nf = eng.register_bank(L.BANK_NOISE, noises)
    return bf, vf, nf

rows = []
def run(name, C, B, mode, aug, seed, metrics=False, T=626):
    bf, vf, nf = banks(C, seed)
    rng = np.random.default_rng(seed)
    if aug:
        d = draw_batch(rng, B, T, bf, vf, nf, max_voices=7, max_noises=2, snr=-20, min_ratio=1, n_time_masks=6, n_freq_masks=1)
    else:
        d = draw_batch(rng, B, T, bf)
    eng.upload_plan(d)
    keep = None
    if aug:
        _, _, keep = eng.labels(); keep = keep.cpu().numpy()
    out = torch.empty(eng.feature_shape(mode), device='cuda')
    bi, bo = eng.plan_bytes(mode, keep)
    y_pred = torch.rand((B, T, 3), device='cuda') if metrics else None
    def step():
        frame = None
        if aug:
            frame, _, _ = eng.labels(want_keep=False)
        eng.features(mode, out=out)
        if metrics:
            eng.metric_counts(frame, y_pred, want_er=False)
    us = timeit(step)
    us_feat = timeit(lambda: eng.features(mode, out=out))
    rows.append('%-58s B=%-5d step %9.1f us  %9.1f kclips/s | features alone %9.1f us  alg %8.1f MB  %6.0f GB/s = %.2f of %.0f' % (
        name, B, us, B / us * 1e3, us_feat, (bi + bo) / 1e6, (bi + bo) / us_feat / 1e3, (bi + bo) / us_feat / 1e3 / PEAK, PEAK))
    print(rows[-1], flush=True)
    del out
    torch.cuda.empty_cache()
    return us

run('configs[0] 2-ch, no augmentation -> min-max log-mel', 2, 32, L.FEAT_LOGMEL_MINMAX, False, 20200)
us1 = run('configs[1] 2-ch mix + masks -> min-max log-mel + labels + counts', 2, 256, L.FEAT_LOGMEL_MINMAX, True, 20201, metrics=True)
run('configs[0] at the trainers\' default n_frame = 512 (sj_train.py:59)', 2, 32, L.FEAT_LOGMEL_MINMAX, False, 20200, T=512)
run('configs[1] at the trainers\' default n_frame = 512 (sj_train.py:59)', 2, 256, L.FEAT_LOGMEL_MINMAX, True, 20201, metrics=True, T=512)
run('configs[2] 4-ch mix + masks -> magnitude + phase + labels', 4, 1024, L.FEAT_MAGPHASE, True, 20202)
run('configs[2] 4-ch mix + masks -> log-magnitude + phase + labels', 4, 1024, L.FEAT_LOG_MAGPHASE, True, 20202)
run('configs[3] shard of 8 GPUs (8192 / 8) -> min-max log-mel + labels + counts', 2, 1024, L.FEAT_LOGMEL_MINMAX, True, 20203, metrics=True)
run('configs[3] whole batch on 1 GPU -> min-max log-mel + labels + counts', 2, 8192, L.FEAT_LOGMEL_MINMAX, True, 20203, metrics=True)

# configs[4]: input-bound check with a SUBSTITUTE consumer of the pipeline's tensors ([B,80,626,2]
# features, [B,626,3] frame labels): a small bf16 conv net (channels-last, 4 stride-2 stages + a
# frame-wise head), forward + backward + SGD step in torch.  Not sj_train's EfficientNet.
B = 256
class Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        ch = [2, 32, 64, 128, 256]
        self.convs = torch.nn.ModuleList([torch.nn.Conv2d(ch[i], ch[i + 1], 3, stride=(2, 1), padding=1) for i in range(4)])
        self.head = torch.nn.Conv1d(256 * 5, 3, 1)
    def forward(self, x):                      # [B, 80, 626, 2]
        x = x.permute(0, 3, 1, 2)              # [B, 2, 80, 626]
        for c in self.convs:
            x = torch.relu(c(x))
        x = x.flatten(1, 2)                    # [B, 256*5, 626]
        return self.head(x).transpose(1, 2)    # [B, 626, 3]
net = Net().cuda().to(memory_format=torch.channels_last)
opt = torch.optim.SGD(net.parameters(), lr=1e-3)
x = torch.randn(B, 80, 626, 2, device='cuda'); y = (torch.rand(B, 626, 3, device='cuda') < 0.3).float()
def train_step():
    with torch.autocast('cuda', dtype=torch.bfloat16):
        loss = torch.nn.functional.binary_cross_entropy_with_logits(net(x).float(), y)
    opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
us_net = timeit(train_step, 5)
rows.append('configs[4] SUBSTITUTE consumer (4-stage bf16 conv net fwd+bwd+SGD, B=256): %.0f us per step; GPU pipeline step %.0f us = %.1f %% of it (not input-bound); the reference CPU pipeline needs ~0.5-0.8 s per 256-clip batch on 16 host cores (bench.py --impl reference: 310-500 clips/s)' % (
    us_net, us1, 100 * us1 / us_net))
print(rows[-1])
open(os.path.join('gpurun_out', 'configs.log'), 'w').write('\n'.join(rows) + '\n')
