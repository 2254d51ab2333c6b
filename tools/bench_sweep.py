"""Inference batch sweep 1 -> 2^20 poses (BASELINE.json configs[1]): device time per p3d_model_forward call
(CUDA events around a run of back-to-back calls on one stream, inputs resident in HBM), poses/s and the
tensor-pipe fraction.  One JSON line per batch size.  Usage: python tools/bench_sweep.py [L nl]"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-pose-baseline_b200")]
from p3d import LinearModel, _lib  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 2
maxlog = int(sys.argv[3]) if len(sys.argv) > 3 else 20
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0}
lib = _lib.lib
m = LinearModel(L, nl, True, True, True, 64, 1e-3, seed=1)
st = torch.cuda.current_stream()
sp = C.c_void_p(st.cuda_stream)
flop_per_pose = 2 * (32 * L + 2 * nl * L * L + L * 48)
for lg in range(0, maxlog + 1):
    B = 1 << lg
    x = torch.randn((B, 32), device="cuda"); y = torch.empty((B, 48), device="cuda")
    iters = 200 if B <= 4096 else (50 if B <= 1 << 16 else 10)
    for _ in range(5):
        _lib.check(lib.p3d_model_forward(m._handle, x.data_ptr(), y.data_ptr(), B, sp))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(iters):
        lib.p3d_model_forward(m._handle, x.data_ptr(), y.data_ptr(), B, sp)
    e1.record(st)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    tf = B * flop_per_pose / (us * 1e-6) / 1e12
    print(json.dumps({"batch": B, "us_per_call": round(us, 2), "poses_per_s": round(B / (us * 1e-6)), "tflops": round(tf, 2),
                      "frac_of_bf16_peak": round(tf / peaks["bf16_tflops"], 4), "linear_size": L, "num_layers": nl}))
m.close()
