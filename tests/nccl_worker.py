"""Worker of tests/test_gpu_step.py::test_nccl_count_allreduce_two_ranks (one process per GPU under
torch.distributed.run).  SURVEY.md 8e: the global batch shards by contiguous clip index; every
rank runs iris_step on its slice with the metric leg + iris_allreduce_counts (NCCL inside libiris,
communicator made by iris_nccl_comm_create); the reduced int64[6] counts and the [B_global,3]
triples must equal the oracle's numbers on the WHOLE batch, and the ER computed from the reduced
triples (global max(n_true), metrics.py:268-273) must equal the oracle's er_score."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from challenge_b200 import _lib as L
    from challenge_b200.dist import shard_range
    from challenge_b200.engine import Engine
    from challenge_b200.plan import draw_config, uniforms_per_clip
    from challenge_b200.synth import synthetic_banks
    from oracle import metrics as M

    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dist.init_process_group('gloo')          # plumbing only: the unique id travels over gloo
    eng = Engine(local)
    eng.set_mel(80)
    bgs, voices, labels, noises = synthetic_banks(99, 2, n_bg=3, n_voice=16, n_noise=4, bg_seconds=4.0)
    eng.register_bank(L.BANK_BG, bgs)
    eng.register_bank(L.BANK_VOICE, voices, labels=labels)
    eng.register_bank(L.BANK_NOISE, noises)
    ids = [eng.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = eng.nccl_comm(ids[0], rank, world)

    Bg, T = 37, 300                           # ragged shards: 19 + 18; T above the longest voice (251 frames)
    lo, hi = shard_range(Bg, world, rank)
    b = hi - lo
    cfg = draw_config(b, T, 5, 2, -20, 1.0, 0.5, 6, 24, 1, 16)
    scfg = eng.step_config(cfg, L.FEAT_LOGMEL_MINMAX)
    rng = np.random.default_rng(1234)
    u_all = rng.random((Bg, uniforms_per_clip(cfg)))
    yp_all = rng.random((Bg, T, 3)).astype(np.float32)
    dev = torch.device('cuda', local)
    y_pred = torch.from_numpy(yp_all[lo:hi]).to(dev)
    counts = torch.zeros(6, dtype=torch.int64, device=dev)
    reduced = torch.zeros(6, dtype=torch.int64, device=dev)
    send = torch.zeros((Bg, 3), dtype=torch.int32, device=dev)   # only rows [lo, hi) are ever written
    glob = torch.zeros((Bg, 3), dtype=torch.int32, device=dev)
    x, frame = eng.step(scfg, u_all[lo:hi], y_pred=y_pred, triples=send[lo:hi], counts=counts, comm=comm,
                        counts_reduced=reduced, triples_send=send, triples_global=glob, global_batch=Bg)
    eng.counts_wait(0)
    torch.cuda.synchronize()
    er = eng.er_from_triples(glob).cpu().numpy()

    # the whole batch's frame labels, gathered over gloo for the check
    frames = [None] * world
    dist.all_gather_object(frames, frame.cpu().numpy())
    yt_all = np.concatenate(frames, 0)
    nt, npd, co = M.er_parts(yt_all, yp_all)
    want_triples = np.stack([nt, npd, co], 1)
    assert np.array_equal(glob.cpu().numpy(), want_triples), 'reduced triples differ from the oracle'
    want6 = np.array(list(M.f1_counts(yt_all, yp_all)) + [nt.sum(), npd.sum(), co.sum()], np.int64)
    assert np.array_equal(reduced.cpu().numpy(), want6), (reduced.cpu().numpy(), want6)
    assert np.array_equal(er, M.er_from_parts(nt, npd, co)), 'ER from the reduced triples differs'
    # the local (unreduced) counts are this rank's share only
    nt_l, np_l, co_l = M.er_parts(yt_all[lo:hi], yp_all[lo:hi])
    assert np.array_equal(counts.cpu().numpy()[3:], [nt_l.sum(), np_l.sum(), co_l.sum()])
    # a per-rank-local ER (clipped at the LOCAL max n_true) would differ when the maxima differ:
    # that is why the triples travel
    dist.barrier()
    eng.nccl_comm_destroy(comm)
    eng.close()
    if rank == 0:
        print('NCCL_WORKER_OK world=%d global_batch=%d counts=%s' % (world, Bg, want6.tolist()), flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
