"""Width-scaled stress variant (BASELINE configs[4]: linear_size=4096, num_layers=4) on one GPU:
    python tools/stress_once.py [log2 poses] [launches]     (P3D_L2_PERSIST=0/1 is read by the library)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-pose-baseline_b200")]
from p3d import LinearModel, _lib  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 21
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B = 1 << lg
model = LinearModel(4096, 4, True, True, True, 64, 1e-3, seed=3, mode="bf16")
g = torch.Generator(device="cuda").manual_seed(21)
x = torch.randn((B, 32), device="cuda", generator=g)
y = torch.empty((B, 48), device="cuda")
for _ in range(2):
    _lib.check(_lib.lib.p3d_model_forward(model._handle, x.data_ptr(), y.data_ptr(), B, None))
torch.cuda.synchronize()
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
clk = []
for _ in range(n):
    _lib.check(_lib.lib.p3d_model_forward(model._handle, x.data_ptr(), y.data_ptr(), B, None))
    clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
e1.record()
while not e1.query():
    clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
ms = e0.elapsed_time(e1) / n
flop = 2 * (32 * 4096 + 2 * 4 * 4096 * 4096 + 4096 * 48)
print(f"P3D_L2_PERSIST={os.environ.get('P3D_L2_PERSIST', 'default')} B=2^{lg}: {ms:.3f} ms/launch, {B / ms / 1e3:.3f} M poses/s, "
      f"{B * flop / ms / 1e9:.1f} TFLOP/s, SM clock min/median {min(clk)}/{sorted(clk)[len(clk) // 2]} MHz, finite={bool(torch.isfinite(y).all())}")
model.close()
