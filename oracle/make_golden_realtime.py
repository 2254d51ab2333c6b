"""Generate tests/golden/realtime.npz by EXECUTING THE REFERENCE's own source lines.

Run in the build container only (needs /root/reference):
    python oracle/make_golden_realtime.py
src/openpose_3dpose_sandbox_realtime.py cannot be imported (TensorFlow, OpenCV, imageio), so the per-frame arithmetic
is cut out of the file BY LINE NUMBER, dedented and exec'ed in a namespace holding the synthetic inputs:
  :69-163  keypoint list handling, BODY_25 surgery, H3.6M scatter, hip/neck/thorax, normalisation
  :178-195 the display shuffle applied to the un-normalised prediction
and data_utils.unNormalizeData is called directly.  Nothing at test/bench time reads /root/reference.
"""
import logging
import os
import sys
import textwrap
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
for n in ["h5py", "matplotlib", "matplotlib.pyplot", "matplotlib.image", "mpl_toolkits", "mpl_toolkits.mplot3d", "viz"]:
    sys.modules[n] = types.ModuleType(n)
sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
sys.path.insert(0, "/root/reference/src")
import data_utils as ref_du          # noqa: E402

SRC = open("/root/reference/src/openpose_3dpose_sandbox_realtime.py").read().split("\n")


def block(lo, hi):
    """Source lines lo..hi (1-based, inclusive), dedented."""
    return textwrap.dedent("\n".join(SRC[lo - 1:hi]))


assert SRC[19].startswith("order = [15, 12, 25"), SRC[19]
assert "_data = data[\"people\"][0][\"pose_keypoints_2d\"]" in SRC[68], SRC[68]
assert "enc_in = np.divide((enc_in - mu), stddev)" in SRC[162], SRC[162]
ORDER = eval(SRC[19].split("=")[1])
FRONT = block(69, 79) + "\n" + block(85, 163)        # skips the regex on the file name (:81-83)
DISPLAY = block(178, 195)
assert "for i in range(poses3d.shape[0]):" in DISPLAY and "spine_x - 630" in DISPLAY

g = np.load(os.path.join(ROOT, "tests", "golden", "geometry.npz"))
m2, s2, use2, ig2 = g["mean2d"], g["std2d"], g["use2d"], g["ignore2d"]
m3, s3, use3, ig3 = g["mean3d"], g["std3d"], g["use3d"], g["ignore3d"]

rng = np.random.RandomState(11)
cases = {"coco54": 18 * 3, "tfpose36": 36, "body25_75": 25 * 3, "wide_87": 29 * 3}
out = {"order": np.array(ORDER), "mean2d": m2, "std2d": s2, "use2d": use2, "ignore2d": ig2,
       "mean3d": m3, "std3d": s3, "use3d": use3, "ignore3d": ig3}
for name, n in cases.items():
    frames = []
    enc_in = np.zeros((1, 64))                       # :28-29, carried from frame to frame like the reference does
    for f in range(3):
        kp = rng.uniform(100, 900, size=n)
        if n % 3 == 0 and n >= 53:
            kp[2::3] = rng.uniform(0, 1, size=n // 3)   # confidence scores
        ns = {"np": np, "order": ORDER, "enc_in": enc_in, "data": {"people": [{"pose_keypoints_2d": kp.tolist()}]},
              "dim_to_use_2d": use2, "data_mean_2d": m2, "data_std_2d": s2,
              "logger": logging.getLogger("x"), "file_name": "f_%d.json" % f, "frame": f}
        exec(FRONT, ns)
        enc_n = ns["enc_in"]                          # [1,32] normalised (:163)
        y = rng.normal(0, 1, size=(1, 48)).astype(np.float32)           # stands in for the model output of :168
        un2d = ref_du.unNormalizeData(enc_n, m2, s2, ig2)                # :170  (next frame's enc_in)
        pose = ref_du.unNormalizeData(y, m3, s3, ig3)                    # :171
        ns2 = {"np": np, "poses3d": pose.copy(), "spine_x": ns["spine_x"], "spine_y": ns["spine_y"]}
        exec(DISPLAY, ns2)
        out["%s_kp%d" % (name, f)] = kp
        out["%s_xy%d" % (name, f)] = np.array(ns["xy"], dtype=np.float64)
        out["%s_enc%d" % (name, f)] = enc_n
        out["%s_spine%d" % (name, f)] = np.array([ns["spine_x"], ns["spine_y"]])
        out["%s_y%d" % (name, f)] = y
        out["%s_pose%d" % (name, f)] = pose
        out["%s_display%d" % (name, f)] = ns2["poses3d"]
        enc_in = un2d
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "realtime.npz"), **out)
print("wrote tests/golden/realtime.npz with", len(out), "arrays")
