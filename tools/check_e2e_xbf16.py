"""Host-buffer step (LinearModel.step with pinned NumPy arrays -> p3d_model_step_eval_host) under P3D_PIPE_XBF16:
the outputs must be BIT-IDENTICAL to the device-resident forward (the host rounds x to bf16 exactly as the pack kernel
does), and the throughput is printed for the A/B.

    python tools/check_e2e_xbf16.py [B]            # default path
    P3D_PIPE_XBF16=1 python tools/check_e2e_xbf16.py [B]
"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-pose-baseline_b200")]
from p3d import LinearModel, _lib  # noqa: E402


def pinned(shape):
    n = int(np.prod(shape)) * 4
    ptr = C.c_void_p()
    _lib.check(_lib.lib.p3d_host_alloc(C.byref(ptr), max(n, 16)))
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_byte)), shape=(max(n, 16),))[:n].view(np.float32).reshape(shape)


B = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 20) + 12345          # ragged last chunk
model = LinearModel(1024, 2, True, True, True, 64, 1e-3, seed=1, mode="bf16")
rng = np.random.RandomState(0)
xh, th, yh = pinned((B, 32)), pinned((B, 48)), pinned((B, 48))
xh[:] = rng.standard_normal((B, 32)).astype(np.float32)
th[:] = rng.standard_normal((B, 48)).astype(np.float32)
xd, td = torch.from_numpy(xh).cuda(), torch.from_numpy(th).cuda()
loss_d, _, yd = model.step(None, xd, td, 1.0, isTraining=False)          # device-resident reference of the same kernels
loss_h, _, y = model.step(None, xh, th, 1.0, isTraining=False, out=yh)
same = bool(np.array_equal(y, yd.cpu().numpy()))
print("P3D_PIPE_XBF16 =", os.environ.get("P3D_PIPE_XBF16", "0"), " B =", B, " outputs bit-identical to the device path:", same,
      " loss host/device: %.7f / %.7f" % (float(loss_h), float(loss_d)))
for tag, t in (("with target", th), ("predictions only", None)):
    model.step(None, xh, t, 1.0, isTraining=False, out=yh)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        model.step(None, xh, t, 1.0, isTraining=False, out=yh)
    torch.cuda.synchronize()
    print("  %s: %.1f M poses/s end to end" % (tag, B * 5 / (time.perf_counter() - t0) / 1e6))
# a small and an odd batch through the same entry point (single-pose chunks keep the fp32 route)
for b in (1, 2, 777, 6144):
    _, _, ys = model.step(None, xh[:b], th[:b], 1.0, isTraining=False)
    _, _, yr = model.step(None, xd[:b], td[:b], 1.0, isTraining=False)
    same = same and bool(np.array_equal(np.asarray(ys), yr.cpu().numpy()))
print("E2E XBF16 CHECK", "OK" if same and abs(float(loss_h) - float(loss_d)) < 1e-6 else "FAIL")
model.close()
sys.exit(0 if same else 1)
