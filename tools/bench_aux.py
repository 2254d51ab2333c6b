"""Secondary measurements (BASELINE.json configs[2] and [3]): camera-frame preprocessing and
Procrustes/MPJPE evaluation against the HBM roofline, the training step, and their CPU baselines
(oracle port on a bounded sample).  One JSON line per workload.  Usage (on the GPU box):
    python tools/bench_aux.py [--cpu]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-pose-baseline_b200")]
from oracle import geometry_ref as G  # noqa: E402
from oracle import synth  # noqa: E402
from p3d import LinearModel, data_utils, evaluate  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle on bounded samples")
ap.add_argument("--no-train", action="store_true", help="skip the training-step timings")
args = ap.parse_args()
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
HBM = peaks["hbm_gbs"]


def timeit(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


N = 1 << 20
cams = synth.cameras(4, seed=3)
g = torch.Generator(device="cuda").manual_seed(0)
root = torch.randn((N, 1, 3), device="cuda", generator=g) * 500
world = (root + torch.randn((N, 32, 3), device="cuda", generator=g) * 300).reshape(N, 96).contiguous()
m2, s2 = np.full(64, 500.0), np.full(64, 150.0)
m3, s3 = np.zeros(96), np.full(96, 200.0)

# ---- fused preprocessing: 2D + 3D, 2D only
ms_both = timeit(lambda: data_utils.camera_frame_dataset(world, cams, m2, s2, m3, s3))
ms_2d = timeit(lambda: data_utils.camera_frame_dataset(world, cams, m2, s2, want_3d=False))
b_both, b_2d = N * (384 + 512 + 768), N * (384 + 512)
for name, ms, b in (("project+normalize 2D+3D, 4 cameras", ms_both, b_both), ("project+normalize 2D, 4 cameras", ms_2d, b_2d)):
    gbs = b / (ms * 1e-3) / 1e9
    print(json.dumps({"workload": name, "poses": N, "ms": ms, "poses_per_s": N / (ms * 1e-3),
                      "roofline": {"bound": "hbm", "achieved": gbs, "peak": HBM, "unit": "GB/s", "frac": gbs / HBM,
                                   "algorithmic_bytes": b}}))

# ---- evaluation: 4 x 2^20 poses (every camera view of the set)
NE = 4 * N
gt = torch.randn((NE, 48), device="cuda", generator=g)
pr = gt + 0.2 * torch.randn((NE, 48), device="cuda", generator=g)
for use_proc in (True, False):
    ms = timeit(lambda: evaluate.mpjpe(pr, gt, m3, s3, procrustes=use_proc), iters=5)
    b = NE * 384
    gbs = b / (ms * 1e-3) / 1e9
    print(json.dumps({"workload": f"un-normalise + {'Procrustes + ' if use_proc else ''}MPJPE", "poses": NE, "ms": ms,
                      "poses_per_s": NE / (ms * 1e-3),
                      "roofline": {"bound": "hbm", "achieved": gbs, "peak": HBM, "unit": "GB/s", "frac": gbs / HBM,
                                   "algorithmic_bytes": b, "note": "includes the D2H read of the 18 sums"}}))

# kernel only (no allocation, no D2H of the sums): what the HBM roofline fraction of the kernel itself is
from p3d import _lib  # noqa: E402
sums = torch.zeros(18, dtype=torch.float64, device="cuda")
m3h, s3h = np.ascontiguousarray(m3, dtype=np.float64), np.ascontiguousarray(s3, dtype=np.float64)
for prec, fn in (("fp32", _lib.lib.p3d_procrustes_mpjpe), ("fp64", _lib.lib.p3d_procrustes_mpjpe_f64)):
    for use_proc in (True, False):
        ms = timeit(lambda: _lib.check(fn(pr.data_ptr(), gt.data_ptr(), _lib.np_ptr(m3h), _lib.np_ptr(s3h), 0, int(use_proc), NE,
                                          None, sums.data_ptr(), _lib.current_stream())), iters=10)
        b = NE * 384
        gbs = b / (ms * 1e-3) / 1e9
        print(json.dumps({"workload": f"kernel only, {prec}: un-normalise + {'Procrustes + ' if use_proc else ''}MPJPE", "poses": NE, "ms": ms,
                          "poses_per_s": NE / (ms * 1e-3),
                          "roofline": {"bound": "hbm", "achieved": gbs, "peak": HBM, "unit": "GB/s", "frac": gbs / HBM,
                                       "algorithmic_bytes": b}}))

# ---- streaming probes: what a plain coalesced stream reaches at each kernel's read:write mix (the measured HBM peak is a 1:1 copy)
src = torch.empty(N * 1664 // 4, dtype=torch.float32, device="cuda").normal_()
dst = torch.empty(N * 1664 // 4, dtype=torch.float32, device="cuda")
for name, rb, wb in (("stream probe 384 B read : 1280 B write per pose (preprocessing 2D+3D mix)", 384, 1280),
                     ("stream probe 384 B read : 512 B write per pose (preprocessing 2D mix)", 384, 512),
                     ("stream probe read only (evaluation mix), 4 x 2^20 x 384 B", 1536, 0),
                     ("stream probe 1:1 copy", 832, 832)):
    ms = timeit(lambda: _lib.check(_lib.lib.p3d_debug_stream_mix(src.data_ptr(), dst.data_ptr(), N * rb // 16, N * wb // 16, _lib.current_stream())))
    gbs = N * (rb + wb) / (ms * 1e-3) / 1e9
    print(json.dumps({"workload": name, "ms": ms, "achieved_gbs": gbs, "frac_of_copy_peak": gbs / HBM}))

# ---- training step (dropout keep 0.5, max_norm, Adam): batch 64 and 4096
for B, mode in (() if args.no_train else ((64, "bf16"), (4096, "bf16"), (64, "fp32"), (4096, "fp32"))):
    model = LinearModel(1024, 2, True, True, True, B, 1e-3, seed=1, mode=mode)
    x = torch.randn((B, 32), device="cuda", generator=g)
    t = torch.randn((B, 48), device="cuda", generator=g)
    ms = timeit(lambda: model.step(None, x, t, 0.5, isTraining=True), warm=3, iters=20)
    flop = B * 25_591_808
    kind = "tcgen05 bf16 GEMMs, fp32 master weights" if mode == "bf16" else "fp32 FFMA GEMMs"
    print(json.dumps({"workload": f"training step batch {B} ({kind})", "ms_per_step": ms, "poses_per_s": B / (ms * 1e-3),
                      "tflops": flop / (ms * 1e-3) / 1e12}))
    model.close()

if args.cpu:
    n = 1 << 15
    w = synth.world_poses(n, seed=5)
    t0 = time.perf_counter()
    G.project_normalize(w, cams, m2, s2, G.dims_to_use(2)[0])
    G.camera_frame_normalize(w, cams, m3, s3, G.dims_to_use(3)[0])
    dt = time.perf_counter() - t0
    print(json.dumps({"cpu_baseline": "project+normalize 2D+3D (NumPy oracle)", "sample_poses": n, "poses_per_s": n / dt, "cores": 1}))
    n = 1 << 16
    gt96, pr96 = synth.eval_pairs(n, seed=4)
    use, ign = G.dims_to_use(3)
    gn = (gt96[:, use] / 200.0).astype(np.float32); pn = (pr96[:, use] / 200.0).astype(np.float32)
    t0 = time.perf_counter()
    G.mpjpe(pn, gn, m3, s3, ign, use, procrustes=True)
    dt = time.perf_counter() - t0
    print(json.dumps({"cpu_baseline": "Procrustes MPJPE (batched NumPy oracle; the reference's per-pose Python loop is ~15 k poses/s)",
                      "sample_poses": n, "poses_per_s": n / dt, "cores": os.cpu_count()}))
