"""End-to-end latency of one realtime frame (keypoints on the host -> un-normalised 3D pose on the host), the frame
loop of src/openpose_3dpose_sandbox_realtime.py:137-171:
  fused   RealtimeLifter.step: one cluster-kernel launch over mapped pinned memory, host spins on a mapped flag
  staged  the same arithmetic as separate calls: NumPy front-end -> LinearModel.step (host buffers) -> p3d unNormalizeData
Host wall-clock per frame (time.perf_counter), p50 / p99 over N frames after warm-up."""
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-pose-baseline_b200")]
from p3d import LinearModel, data_utils  # noqa: E402
from p3d.realtime import ORDER, RealtimeLifter  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
g = np.load(os.path.join(ROOT, "tests", "golden", "realtime.npz"))
m2, s2, use2, ig2, m3, s3, use3, ig3 = (g[k] for k in ["mean2d", "std2d", "use2d", "ignore2d", "mean3d", "std3d", "use3d", "ignore3d"])
model = LinearModel(1024, 2, True, True, True, 64, 1e-3, seed=1)
lifter = RealtimeLifter(model, m2, s2, use2, m3, s3, use3)
rng = np.random.RandomState(0)
frames = rng.uniform(100, 900, size=(64, 36))


def staged(xy):
    enc = np.zeros((1, 64))
    for i, j in enumerate(ORDER):
        enc[0, j * 2:j * 2 + 2] = xy[i * 2:i * 2 + 2]
    enc[0, 0:2] = (enc[0, 2:4] + enc[0, 12:14]) / 2
    enc[0, 28:30] = (enc[0, 30:32] + enc[0, 24:26]) / 2
    enc[0, 26:28] = 2 * enc[0, 24:26] - enc[0, 28:30]
    e = (enc[:, use2] - m2[use2]) / s2[use2]
    _, _, y = model.step(None, e, np.zeros((1, 48)), 1.0, isTraining=False)
    return data_utils.unNormalizeData(y, m3, s3, ig3)


def timeit(fn):
    for i in range(300):
        fn(frames[i % 64])
    t = []
    for i in range(N):
        t0 = time.perf_counter()
        fn(frames[i % 64])
        t.append((time.perf_counter() - t0) * 1e6)
    t.sort()
    return {"p50_us": round(statistics.median(t), 2), "p99_us": round(t[int(0.99 * len(t))], 2), "min_us": round(t[0], 2)}


a = lifter.step(frames[0])[2]
b = staged(frames[0])
res = {"workload": "one realtime frame, host keypoints -> host 3D pose (linear_size 1024, bf16)", "frames": N,
       "fused_one_launch": timeit(lambda xy: lifter.step(xy)), "staged_calls": timeit(staged),
       "max_abs_diff_mm": float(np.abs(a - b).max())}
print(json.dumps(res))
lifter.close()
model.close()
