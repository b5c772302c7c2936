"""Multi-GPU plumbing: the batch shards by contiguous clip index, one process per GPU; the
path's only exchange is the sum all-reduce of integer metric counts.

Because the reference's ER is a mean of per-sample ratios whose denominator clips at the
batch-global ``max(n_true)`` (metrics.py:271-273), exact parity of the reduced metric needs
the per-sample triples of every rank: each rank fills its slice of a zero-initialised
``[B_global, 3]`` buffer and the buffers are summed (sum over disjoint slices == all-gather)
in the same collective as the ``[TP, FP, FN]`` vector.
"""
import numpy as np


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPU cores NVML reports as local to GPU ``device_index`` -- call it
    before the first pinned allocation: one process per GPU, each on its GPU's NUMA node, so that
    the pinned staging buffers of the plan uploads and of the feature read-back are node-local
    and the ranks of one box do not all write into one socket's memory.  Returns the core list,
    or None when NVML / the affinity call is not available (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        idx = int(device_index)
        if vis:
            ids = [v.strip() for v in vis.split(',') if v.strip()]
            if idx < len(ids) and ids[idx].isdigit():
                idx = int(ids[idx])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cores = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cores = sorted(c for c in cores if c in allowed)
        if not cores:
            return None
        os.sched_setaffinity(0, cores)
        return cores
    except Exception:
        return None


def shard_range(global_batch, world_size, rank):
    """Contiguous clip range ``[lo, hi)`` of ``rank`` (remainder spread over the first ranks)."""
    base, rem = divmod(int(global_batch), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_counts(tpfpfn, triples, global_batch, lo):
    """One int64 vector ``[TP, FP, FN, triples of all B_global samples...]`` for a single
    all-reduce; ``triples`` is this rank's ``[b_local, 3]`` slice starting at clip ``lo``."""
    import torch
    buf = torch.zeros(3 + 3 * int(global_batch), dtype=torch.int64, device=triples.device)
    buf[:3] = tpfpfn.to(torch.int64)
    n = triples.shape[0]
    buf[3 + 3 * lo:3 + 3 * (lo + n)] = triples.reshape(-1).to(torch.int64)
    return buf


def allreduce_counts(buf, group=None):
    """Sum over ranks (NCCL on GPU tensors over NVLink; gloo on CPU tensors in tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


def unpack_counts(buf):
    """-> (tpfpfn [3], triples [B_global, 3])."""
    return buf[:3], buf[3:].reshape(-1, 3)


def er_from_triples(triples):
    """metrics.py:268-273 on the GLOBAL triples: fp32 score per sample."""
    t = np.asarray(triples.cpu() if hasattr(triples, 'cpu') else triples)
    f32 = np.float32
    n_true = t[:, 0].astype(f32)
    score = n_true + t[:, 1].astype(f32) - f32(2) * t[:, 2].astype(f32)
    hi = n_true.max() if n_true.size else f32(0)
    with np.errstate(divide='ignore', invalid='ignore'):
        return (score / np.minimum(np.maximum(n_true, f32(1)), hi)).astype(f32)


def f1_from_counts(tp, fp, fn):
    """tfa F1Score(average='micro').result(): divide_no_nan everywhere (metrics.py:291)."""
    f32 = np.float32

    def dnn(a, b):
        return f32(0) if b == 0 else f32(a) / f32(b)

    tp, fp, fn = f32(tp), f32(fp), f32(fn)
    p, r = dnn(tp, tp + fp), dnn(tp, tp + fn)
    return f32(dnn(p * r, p + r) * f32(2))
