#!/usr/bin/env python
"""Benchmark of the hot path: augmented spectrogram clips/sec (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # the CUDA path (libiris)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm

Workload (`config.workload`): BASELINE.json configs[1] -- synthetic 2-ch 16 kHz 10 s clips,
background + voices + noise mixed at sampled gains, 6 time + 1 frequency mask, batch 256,
n_fft 512 / hop 256 / 80 mel, min-max + log, frame labels -- on every GPU (weak scaling:
each rank draws and produces its own 256-clip slice), followed by the integer F1 / error-rate
counting of the frame labels against a synthetic prediction; for N > 1 the count vector is
all-reduced over NCCL (the path's only exchange).

One step = one batch.  `value` is timed with CUDA events on the launching stream with the
plan already resident (L2 flushed between steps); `e2e` goes through the public drop-in
call with host draws in, pinned host features + labels out.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# The CPU arm runs one process per core: BLAS / OpenMP pools inside each worker would
# oversubscribe the box (round 1: 82 clips/s on 16 cores, ~500 with one thread per worker).
# numpy reads these when it is first imported, so they are set before that; the GPU arm only
# uses numpy for the host draws.
_THREAD_ENV = ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS', 'NUMEXPR_NUM_THREADS',
               'VECLIB_MAXIMUM_THREADS')
for _k in _THREAD_ENV:
    os.environ[_k] = '1'

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(n_chan=2, batch=256, n_frame=626, max_voices=7, max_noises=2, snr=-20, min_ratio=1,
           n_mels=80, n_time_masks=6, n_freq_masks=1, n_bg=64, n_voice=256, n_noise=64,
           seed=20202)
# dram__bytes_read.sum + dram__bytes_write.sum of one k_fused<FM_MEL> launch of this workload,
# from the committed ncu --set full capture (profiles/); below the algorithmic bytes because the
# 175 MB of banks are shared by the 256 clips and partly stay in the 126 MB L2
NCU_TRAFFIC_BYTES = 439142144
NCU_TRAFFIC_SRC = 'profiles/r01_v10_fused_ncu_raw.txt (411.4 MB read + 27.8 MB written; the mel rows stay in L2 for k_logmel_post)'
METRIC = 'augmented spectrogram clips/sec'
UNIT = 'clips/s'


def workload_config(n_gpus):
    return {
        'workload': 'BASELINE configs[1]: 2-ch 16 kHz 10 s clips + noise mixing at random SNR '
                    '+ time/freq masking -> min-max log-mel [B,80,626,2] + frame labels '
                    '[B,626,3] + F1/ER counts',
        'batch_per_gpu': CFG['batch'], 'global_batch': CFG['batch'] * n_gpus,
        'n_fft': 512, 'hop': 256, 'n_mels': CFG['n_mels'], 'n_frame': CFG['n_frame'],
        'max_voices': CFG['max_voices'], 'max_noises': CFG['max_noises'],
        'banks': '%d bg x 10 s, %d voices / %d noises 0.5-4 s (seed %d)' % (
            CFG['n_bg'], CFG['n_voice'], CFG['n_noise'], CFG['seed']),
        'parallelism': 'dp%d (batch sharded by clip, int64 count all-reduce only)' % n_gpus,
        'l2': 'flushed between timed steps (256 MiB write); banks + outputs exceed L2',
    }


# --------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': float(np.max(mx)) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback 6.65 TB/s (B200_PROFILING.md)'


# --------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference pipeline, timed on host cores
_CPU = {}


def _cpu_init(seed):
    import torch
    torch.set_num_threads(1)
    from challenge_b200.synth import synthetic_banks
    _CPU['raw'] = synthetic_banks(seed, CFG['n_chan'], CFG['n_bg'], CFG['n_voice'], CFG['n_noise'])
    _CPU['spec'] = ({}, {}, {})


class _LazyBank:
    """load_wav of a source on first use (the reference does this offline; excluded from
    the timed region by warming the cache before timing)."""

    def __init__(self, kind):
        self.kind = kind

    def __getitem__(self, i):
        from oracle.data_utils import load_wav_array
        cache = _CPU['spec'][self.kind]
        if i not in cache:
            raw = _CPU['raw'][(0, 1, 3)[self.kind]]
            cache[i] = load_wav_array(raw[i])
        return cache[i]


def _cpu_clips(args):
    """Run the reference chain for clips [lo, hi) of draws d; returns seconds (timed part)."""
    d, lo, hi, warm = args
    from oracle import chain
    banks = (_LazyBank(0), _LazyBank(1), _LazyBank(2))
    if warm:   # offline part: STFT of every source these clips touch
        for b in range(lo, hi):
            banks[0][int(d.bg_id[b])]
            for i in d.voice_id[b]:
                banks[1][int(i)]
            for i in d.noise_id[b]:
                banks[2][int(i)]
        return 0.0
    t0 = time.perf_counter()
    x, y, _, _ = chain.dataset_batch(banks[0], banks[1], _CPU['raw'][2], banks[2], d,
                                     mode='logmel_minmax', n_mels=CFG['n_mels'],
                                     clips=range(lo, hi))
    from oracle import metrics as M
    yp = np.clip(y + 0.25, 0, 1).astype(np.float32)
    M.er_parts(y, yp)
    M.f1_counts(y, yp)
    return time.perf_counter() - t0


def cpu_draws(n_clips, seed):
    from challenge_b200.plan import draw_batch
    raw = _CPU['raw']
    frames = [np.array([1 + w.shape[1] // 256 for w in raw[k]], np.int32) for k in (0, 1, 3)]
    rng = np.random.default_rng(seed)
    return draw_batch(rng, n_clips, CFG['n_frame'], frames[0], frames[1], frames[2],
                      max_voices=CFG['max_voices'], max_noises=CFG['max_noises'], snr=CFG['snr'],
                      min_ratio=CFG['min_ratio'], n_time_masks=CFG['n_time_masks'],
                      n_freq_masks=CFG['n_freq_masks'])


class CpuPool:
    """All host cores, one process per core over clips (fork; banks shared copy-on-write)."""

    def __init__(self, cores):
        import multiprocessing as mp
        self.cores = cores
        _cpu_init(CFG['seed'])
        self.pool = mp.get_context('fork').Pool(cores, initializer=_cpu_init,
                                                initargs=(CFG['seed'],)) if cores > 1 else None

    def run(self, d, n_clips, warm=False):
        """Wall seconds for n_clips clips of d spread over the cores."""
        per = -(-n_clips // self.cores)
        chunks = [(d, lo, min(lo + per, n_clips), warm) for lo in range(0, n_clips, per)]
        t0 = time.perf_counter()
        if self.pool is None:
            for c in chunks:
                _cpu_clips(c)
        else:
            self.pool.map(_cpu_clips, chunks)
        return time.perf_counter() - t0

    def close(self):
        if self.pool is not None:
            self.pool.terminate()


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference_arm(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on all host cores;
    each step is a bounded sample of the same workload."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = host_cores()
    per_step = max(2 * cores, 8)
    pool = CpuPool(cores)
    try:
        d = cpu_draws(per_step * (args.steps + args.warmup), CFG['seed'] + 1)
        pool.run(d, d.batch, warm=True)
        times = []
        for s in range(args.warmup + args.steps):
            ds = d.slice(s * per_step, (s + 1) * per_step)
            dt = pool.run(ds, per_step)
            if s >= args.warmup:
                times.append(dt)
    finally:
        pool.close()
    total = float(np.sum(times))
    value = per_step * args.steps / total
    sample = ('%d clips per step on %d processes (1 torch thread each); per-source load_wav '
              '(STFT) precomputed offline as in the reference' % (per_step, cores))
    out = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.gpus),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from challenge_b200 import _lib as L
    from challenge_b200.engine import Engine
    from challenge_b200.plan import draw_batch
    from challenge_b200.synth import synthetic_banks

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    from challenge_b200.dist import bind_to_gpu_numa
    numa_cores = bind_to_gpu_numa(local) if world > 1 else None   # before any pinned allocation
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    eng = Engine(local)
    eng.set_mel(CFG['n_mels'])
    bgs, voices, labels, noises = synthetic_banks(CFG['seed'], CFG['n_chan'], CFG['n_bg'],
                                                  CFG['n_voice'], CFG['n_noise'])
    bf = eng.register_bank(L.BANK_BG, bgs)
    vf = eng.register_bank(L.BANK_VOICE, voices, labels=labels)
    nf = eng.register_bank(L.BANK_NOISE, noises)
    B, T, K = CFG['batch'], CFG['n_frame'], 3
    rng = np.random.default_rng(CFG['seed'] + 100 + rank)

    def draw():
        return draw_batch(rng, B, T, bf, vf, nf, max_voices=CFG['max_voices'],
                          max_noises=CFG['max_noises'], snr=CFG['snr'], min_ratio=CFG['min_ratio'],
                          n_time_masks=CFG['n_time_masks'], n_freq_masks=CFG['n_freq_masks'])

    n_all = args.warmup + args.steps
    feat = [torch.empty((B, CFG['n_mels'], T, CFG['n_chan']), device=dev) for _ in range(2)]
    # stand-in for the model output the metric kernels score (fixed, generated once)
    y_pred = torch.rand((B, T, K), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # one int64 [TP, FP, FN, sum n_true, sum n_pred, sum correct] row per step: filled by
    # k_metric_counts, all-reduced in place over NCCL (the path's only exchange)
    n_e2e_all = 0 if args.no_e2e else args.steps + max(args.warmup, 3)
    counts = torch.zeros((n_all + n_e2e_all + 1, 6), dtype=torch.int64, device=dev)
    KERNELS = ['k_labels', 'k_tiles', 'k_fused<FM_MEL>', 'k_logmel_post', 'k_metric_counts']

    side = torch.cuda.Stream(device=dev)
    ev_lab = torch.cuda.Event()
    ev_met = [torch.cuda.Event() for _ in range(2)]
    met_pending = [False, False]

    def device_step(i, out):
        """labels -> {fused features  ||  metric counts (+ count all-reduce)}: the counting only
        needs the frame labels, so it runs on a second stream beside the feature kernel.  On one
        GPU the step ends when both are done.  With several ranks the count all-reduce is also a
        rendezvous of the ranks, so a step waits for the collective of the step BEFORE it (the
        reduced counts are consumed one step late, like a logged Keras metric); `drain_counts`
        waits for the last one inside the timed region."""
        main = torch.cuda.current_stream()
        frame, _, _ = eng.labels(want_keep=False)
        ev_lab.record(main)
        j = i & 1
        with torch.cuda.stream(side):
            side.wait_event(ev_lab)
            eng.metric_counts(frame, y_pred, counts=counts[i], want_er=False)
            if world > 1:
                dist.all_reduce(counts[i])                  # NCCL sum of the int64 count vector
            ev_met[j].record(side)
            met_pending[j] = True
            frame.record_stream(side)
        eng.features(L.FEAT_LOGMEL_MINMAX, out=out)
        k = j if world == 1 else j ^ 1
        if met_pending[k]:
            main.wait_event(ev_met[k])
            met_pending[k] = False
        return frame, ev_met[j]

    def drain_counts():
        main = torch.cuda.current_stream()
        for k in range(2):
            if met_pending[k]:
                main.wait_event(ev_met[k])
                met_pending[k] = False

    # ---- kernel-resident timing: plan already uploaded, CUDA events per step ----
    plans = [draw() for _ in range(n_all)]
    eng.profile(False)
    for s in range(args.warmup):
        eng.upload_plan(plans[s])
        device_step(s, feat[0])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    eng.profile(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    for s in range(args.steps):
        eng.upload_plan(plans[args.warmup + s])
        flush.fill_(s & 0xff)                               # evict L2 (untimed)
        ev[s][0].record()
        device_step(args.warmup + s, feat[0])
        if s == args.steps - 1:
            drain_counts()                                  # the last collective ends inside the timed region
        ev[s][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(np.sum(step_ms))
    fused_ms, n_fused = eng.profile_read()
    eng.profile(False)
    # algorithmic bytes of the last plan (kept sources only), as an average per step
    _, _, keep = eng.labels()
    bi, bo = eng.plan_bytes(L.FEAT_LOGMEL_MINMAX, keep.cpu().numpy())
    alg_bytes = bi + bo
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * B * args.steps / (total_ms_max / 1e3)

    # ---- end to end through the public call: host draws in, pinned host tensors out; the
    # device->host copies of step i overlap the kernels of step i+1 (two buffers, two streams) ----
    h_feat = [torch.empty(feat[0].shape, dtype=torch.float32, pin_memory=True) for _ in range(2)]
    h_lbl = [torch.empty((B, T, K), dtype=torch.float32, pin_memory=True) for _ in range(2)]
    h_cnt = [torch.empty(6, dtype=torch.int64, pin_memory=True) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    done = [torch.cuda.Event() for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    h2d = d2h = 0

    # host randomness (numpy, ~0.4 ms per 256-clip batch) is drawn one step ahead on a worker
    # thread, like a tf.data prefetch(1): the draws of step i+1 overlap the launches of step i
    from concurrent.futures import ThreadPoolExecutor
    drawer = ThreadPoolExecutor(max_workers=1)
    next_draw = [drawer.submit(draw)]

    def e2e_step(i, features_to_host=True):
        nonlocal h2d, d2h
        j = i & 1
        row = n_all + i
        d = next_draw[0].result()                           # host randomness (numpy), drawn ahead
        next_draw[0] = drawer.submit(draw)
        torch.cuda.current_stream().wait_event(copied[j])   # buffer j is free again
        info = eng.upload_plan(d)                           # H2D of the draws (pinned staging)
        frame, counted = device_step(row, feat[j])
        done[j].record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done[j])
            copy_stream.wait_event(counted)                 # this step's (all-reduced) counts
            if features_to_host:
                h_feat[j].copy_(feat[j], non_blocking=True)
            h_lbl[j].copy_(frame, non_blocking=True)
            h_cnt[j].copy_(counts[row], non_blocking=True)
            copied[j].record()
            frame.record_stream(copy_stream)
        if features_to_host:
            h2d = info['bytes']
            d2h = h_feat[j].numel() * 4 + h_lbl[j].numel() * 4 + h_cnt[j].numel() * 8

    n_e2e = 0 if args.no_e2e else args.steps
    n_e2e_warm = 0 if args.no_e2e else max(args.warmup, 3)
    for i in range(n_e2e_warm):
        e2e_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(n_e2e):
        e2e_step(n_e2e_warm + i)
    torch.cuda.synchronize()
    e2e_s = max(time.perf_counter() - t0, 1e-9)
    clocks = sampler.stop()
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * n_e2e / float(t.item())
    # the same loop with the features left on the device (what a Keras / torch model fed through
    # DLPack sees: INTEGRATION.md section 3); labels and counts still go to the host.  Context for
    # the PCIe-bound number above, not the e2e value.
    for i in range(min(n_e2e_warm, 2)):
        e2e_step(i, features_to_host=False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(n_e2e):
        e2e_step(i, features_to_host=False)
    torch.cuda.synchronize()
    t = torch.tensor([max(time.perf_counter() - t0, 1e-9)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_dev_value = world * B * n_e2e / float(t.item())
    next_draw[0].result()
    drawer.shutdown()

    peak, peak_src = measured_peak()
    fused_avg_ms = fused_ms / max(n_fused, 1)
    achieved = alg_bytes / (fused_avg_ms / 1e3) / 1e9 if fused_avg_ms > 0 else 0.0
    out = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': total_ms_max / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': workload_config(world),
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
                'd2h_bytes_per_step': int(d2h),
                'note': 'host draws (numpy, one step ahead on a worker thread) -> iris_plan_upload (H2D) -> kernels -> features + '
                        'labels + counts copied to pinned host memory; wall clock; the D2H of '
                        'step i overlaps the kernels of step i+1; PCIe-bound (features are '
                        '400 KB per clip)',
                'features_on_device_value': e2e_dev_value,
                'features_on_device_note': 'same loop, features handed over on the device (DLPack) '
                                           'instead of copied to the host; labels + counts still read back'},
        'kernels_per_step': KERNELS,
        'roofline': {'bound': 'hbm', 'kernel': 'k_fused<FM_MEL> (+ k_tiles)', 'achieved': achieved,
                     'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak if peak else None,
                     'traffic': NCU_TRAFFIC_BYTES, 'traffic_source': NCU_TRAFFIC_SRC,
                     'peak_source': peak_src,
                     'algorithmic_bytes_per_launch': int(alg_bytes),
                     'kernel_ms': fused_avg_ms, 'kernel_share_of_step': fused_ms / total_ms},
    }
    if world > 1:
        out['config']['host_binding'] = ('each rank pinned to the %d cores NVML lists for its GPU' % len(numa_cores)
                                         if numa_cores else 'none (NVML affinity not available)')
        out['config']['count_allreduce'] = 'one NCCL all-reduce of int64[6] per step on a side stream; a step waits for the collective of the previous step, the last one is drained inside the timed region'
    out['gpu_launches'] = int(args.steps * (len(KERNELS) + (1 if world > 1 else 0)))
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:   # the CPU leg is timed at N = 1 only
            out['cpu_baseline'] = cpu_baseline()
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline():
    """The oracle port of the reference pipeline timed on this box's host cores on a bounded
    sample of the same workload (reported beside the GPU number; not the target)."""
    cores = host_cores()
    n = max(4 * cores, 16)
    pool = CpuPool(cores)
    try:
        d = cpu_draws(n, CFG['seed'] + 2)
        pool.run(d, n, warm=True)
        wall = pool.run(d, n)
        one = CpuPool(1)
        n1 = 8
        one.run(d.slice(0, n1), n1, warm=True)
        t1 = one.run(d.slice(0, n1), n1)
    finally:
        pool.close()
    return {'value': n / wall, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'single_thread_value': n1 / t1,
            'sample': '%d clips of the same workload on %d processes (1 torch thread each), '
                      'per-source load_wav (STFT) precomputed offline as in the reference; '
                      'single_thread_value: %d clips on 1 thread' % (n, cores, n1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true', help='profiling runs only')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == '__main__':
    main()
