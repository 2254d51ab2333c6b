"""world_size-2 gloo tests (CPU) of the data-parallel host protocol: contiguous row shards, the
all-reduce of per-joint error sums + pose count (evaluate.mpjpe) and of BN (sum, sumsq) statistics
(SyncBN) reproduce the single-process result.  The oracle stands in for the device compute."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "3d-pose-baseline_b200")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import geometry_ref as G, synth
    from p3d.linear_model import shard_rows
    N = 101
    gt96, pr96 = synth.eval_pairs(N, seed=4)
    use, ign = G.dims_to_use(3)
    mean = np.zeros(96)
    std = np.full(96, 150.0)
    gt_n = (gt96[:, use] - mean[use]) / std[use]
    pr_n = ((pr96[:, use] - mean[use]) / std[use]).astype(np.float32)
    lo, hi = shard_rows(N, rank, world)
    d = G.mpjpe(pr_n[lo:hi], gt_n[lo:hi], mean, std, ign, use, procrustes=True)
    sums = torch.from_numpy(np.concatenate([d.sum(0), [hi - lo]]))
    dist.all_reduce(sums)
    # SyncBN statistics
    z = np.random.RandomState(1).standard_normal((N, 8))
    st = torch.from_numpy(np.stack([z[lo:hi].sum(0), (z[lo:hi] ** 2).sum(0)]))
    dist.all_reduce(st)
    if rank == 0:
        full = G.mpjpe(pr_n, gt_n, mean, std, ign, use, procrustes=True)
        s = sums.numpy()
        ok1 = np.allclose(s[:17] / s[17], full.mean(0), rtol=1e-12) and s[17] == N
        m = st.numpy()[0] / N
        v = st.numpy()[1] / N - m ** 2
        ok2 = np.allclose(m, z.mean(0)) and np.allclose(v, z.var(0))
        q.put(bool(ok1 and ok2))
    dist.destroy_process_group()


def test_sharded_reductions_match_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
    assert ok
