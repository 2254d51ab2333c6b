"""Batch-1 (and small batch) latency of LinearModel inference: CUDA-event p50 per call.  Env knobs of the
latency kernel (P3D_LAT_GRID / P3D_LAT_THREADS / P3D_LAT_COOP) are read by the library at first use."""
import ctypes as C
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-pose-baseline_b200")]
from p3d import LinearModel, _lib  # noqa: E402

lib = _lib.lib
m = LinearModel(1024, 2, True, True, True, 64, 1e-3, seed=1)
st = torch.cuda.current_stream()
sp = C.c_void_p(st.cuda_stream)
for B in [int(b) for b in (sys.argv[1:] or ["1", "8", "16"])]:
    x = torch.randn((B, 32), device="cuda"); y = torch.empty((B, 48), device="cuda")
    for _ in range(200):
        _lib.check(lib.p3d_model_forward(m._handle, x.data_ptr(), y.data_ptr(), B, sp))
    torch.cuda.synchronize()
    ev, wall = [], []
    for _ in range(1000):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        a.record(st)
        _lib.check(lib.p3d_model_forward(m._handle, x.data_ptr(), y.data_ptr(), B, sp))
        b.record(st)
        b.synchronize()
        wall.append((time.perf_counter() - w0) * 1e6)
        ev.append(a.elapsed_time(b) * 1e3)
    # back-to-back throughput (launch pipelining)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(1000):
        lib.p3d_model_forward(m._handle, x.data_ptr(), y.data_ptr(), B, sp)
    torch.cuda.synchronize(); bb = (time.perf_counter() - t0) * 1e3
    print(f"B={B}: p50 device {statistics.median(ev):.1f} us, p50 wall {statistics.median(wall):.1f} us, back-to-back {bb:.1f} us/call "
          f"[grid={os.environ.get('P3D_LAT_GRID','auto')} threads={os.environ.get('P3D_LAT_THREADS','1024')} coop={os.environ.get('P3D_LAT_COOP','1')}]")
if os.environ.get("P3D_LAT_STAMPS"):
    import numpy as np
    st_ = np.zeros(16, np.uint64)
    _lib.check(lib.p3d_debug_latency_stamps(m._handle, st_.ctypes.data, 7))
    d = (st_[1:7].astype(np.int64) - st_[0:6].astype(np.int64))
    print("in-kernel phase times (ns): layer0+setup, hidden1..4, output:", d.tolist(), "total", int(st_[6] - st_[0]))
m.close()
