"""A few launches of one tc_gemm problem (for ncu captures): python tools/gemm_once.py M N K [a_mn b_mn colsum res split_k iters]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-pose-baseline_b200")]
from p3d import _lib  # noqa: E402

a = [int(v) for v in sys.argv[1:]] + [0] * 9
M, N, K, a_mn, b_mn, colsum, res, split_k, iters = a[:9]
iters = iters or 6
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn((K, M) if a_mn else (M, K), generator=g, device="cuda").to(torch.bfloat16)
B = torch.randn((K, N) if b_mn else (N, K), generator=g, device="cuda").to(torch.bfloat16)
C = torch.zeros((M, N), dtype=torch.float32, device="cuda")
bias = torch.randn(N, generator=g, device="cuda")
rv = torch.randn((M, N), generator=g, device="cuda") if res else None
cs = torch.zeros(2 * N, dtype=torch.float64, device="cuda") if colsum else None
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(iters):
    if i == 2:
        e0.record()
    _lib.check(_lib.lib.p3d_debug_tc_gemm(A.data_ptr(), A.shape[1], a_mn, B.data_ptr(), B.shape[1], b_mn, C.data_ptr(), N, M, N, K,
                                          bias.data_ptr(), rv.data_ptr() if res else None, 1.0, split_k,
                                          cs.data_ptr() if colsum else None, None))
e1.record()
torch.cuda.synchronize()
print(f"M={M} N={N} K={K}: {e0.elapsed_time(e1) / (iters - 2) * 1e3:.1f} us/launch, checksum {float(C.sum()):.3e}")
