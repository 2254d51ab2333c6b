"""tcgen05.mma issue/completion cost as a function of N (128 x N x 16, bf16, operands resident in smem)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-pose-baseline_b200")]
from p3d import _lib
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for N in (16, 32, 64, 128, 256):
    for iters in (4, 64, 256):
        _lib.check(_lib.lib.p3d_debug_mma_rate(N, iters, out.data_ptr(), None)); torch.cuda.synchronize()
        _lib.check(_lib.lib.p3d_debug_mma_rate(N, iters, out.data_ptr(), None)); torch.cuda.synchronize()
        o = out.cpu().tolist()
        print(f"N={N:3d} iters={iters:3d}: issue {o[0]:6d} cyc ({o[0]/iters:6.1f}/mma)  complete {o[1]:6d} cyc ({o[1]/iters:6.1f}/mma)")
