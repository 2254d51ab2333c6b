"""A few device-resident launches of the fused inference kernel (for ncu): python tools/forward_once.py [B] [launches]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-pose-baseline_b200")]
from p3d import LinearModel, _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
model = LinearModel(1024, 2, True, True, True, 64, 1e-3, seed=1, mode="bf16")
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((B, 32), device="cuda", generator=g)
y = torch.empty((B, 48), device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
_lib.check(_lib.lib.p3d_model_forward(model._handle, x.data_ptr(), y.data_ptr(), B, None))
torch.cuda.synchronize()
e0.record()
for _ in range(n):
    _lib.check(_lib.lib.p3d_model_forward(model._handle, x.data_ptr(), y.data_ptr(), B, None))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"P3D_L2_PERSIST={os.environ.get('P3D_L2_PERSIST', '0')} B={B}: {ms:.3f} ms/launch, {B / ms / 1e3:.1f} M poses/s, finite={bool(torch.isfinite(y).all())}")
model.close()
