"""Data-parallel training step time (launch with torchrun, one process per GPU): the GLOBAL batch is split by rows,
SyncBN sums and the flat gradient go through NCCL.  CUDA events, max over ranks.  One JSON line per global batch."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-pose-baseline_b200")]
from p3d import LinearModel  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
for B in [int(b) for b in (sys.argv[1:] or [64, 4096, 32768])]:
    for mode in ("bf16",):
        m = LinearModel(1024, 2, True, True, True, B, 1e-3, mode=mode, device=local, seed=1, dist=dist if world > 1 else None)
        g = torch.Generator(device=dev).manual_seed(0)
        x = torch.randn((B, 32), device=dev, generator=g); t = torch.randn((B, 48), device=dev, generator=g)
        for _ in range(5):
            m.step(None, x, t, 0.5, isTraining=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for _ in range(iters):
            loss, _, _, _ = m.step(None, x, t, 0.5, isTraining=True)
        e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"workload": "data-parallel training step (dropout 0.5, max_norm, Adam, SyncBN)", "n_gpus": world, "global_batch": B,
                              "rows_per_gpu": B // world, "mode": mode, "us_per_step": round(float(ms.item()) * 1e3, 1),
                              "poses_per_s": round(B / (float(ms.item()) * 1e-3)), "loss": round(float(loss), 4), "peer_memory_reductions": bool(getattr(m, "p2p", False))}))
        m.close()
if world > 1:
    dist.destroy_process_group()
