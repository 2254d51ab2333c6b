"""Run a few training steps (for ncu launch lists / timing): python tools/train_steps.py B [mode] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-pose-baseline_b200")]
from p3d import LinearModel  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
model = LinearModel(1024, 2, True, True, True, B, 1e-3, seed=1, mode=mode)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((B, 32), device="cuda", generator=g)
t = torch.randn((B, 48), device="cuda", generator=g)
for _ in range(steps):
    loss, _, _, _ = model.step(None, x, t, 0.5, isTraining=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss, _, _, _ = model.step(None, x, t, 0.5, isTraining=True)
e1.record(); torch.cuda.synchronize()
print(f"B={B} mode={mode}: {e0.elapsed_time(e1) / steps * 1e3:.1f} us/step, loss {float(loss):.4f}")
model.close()
