"""Generate tests/golden/driver.npz by EXECUTING THE REFERENCE's own source lines for the two host-side
callers of the hot path.

Run in the build container only (needs /root/reference):
    python oracle/make_golden_driver.py
Neither file can be imported (both `import tensorflow` at the top), so - as oracle/make_golden_realtime.py does for
the realtime script - the functions are cut out BY LINE NUMBER, dedented and exec'ed in a namespace that holds the
reference's own data_utils / procrustes modules and stand-ins for what TensorFlow would have provided:
  src/linear_model.py:247-300    LinearModel.get_all_batches (pure NumPy; `self` = an object with input_size,
                                 output_size, batch_size)
  src/predict_3dpose.py:352-444  evaluate_batches (the un-normalise / Procrustes / per-joint error loop); `model.step`
                                 is a stub that hands back recorded predictions and per-batch losses (the part that
                                 is TensorFlow's), FLAGS an object with batch_size / procrustes / predict_14
Nothing at test/bench time reads /root/reference.
"""
import os
import sys
import textwrap
import time
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
import procrustes as ref_pr          # noqa: E402

from oracle import synth             # noqa: E402


def block(path, lo, hi):
    src = open(path).read().split("\n")
    return src, textwrap.dedent("\n".join(src[lo - 1:hi]))


# ------------------------------------------------------------------ LinearModel.get_all_batches
src, GAB = block("/root/reference/src/linear_model.py", 247, 300)
assert src[246].strip().startswith("def get_all_batches( self, data_x, data_y, camera_frame, training=True ):"), src[246]
assert src[299].strip() == "return encoder_inputs, decoder_outputs", src[299]
ns = {"np": np}
exec(GAB, ns)
ref_get_all_batches = ns["get_all_batches"]

out = {}
rng = np.random.RandomState(21)
# keys as train() builds them (predict_3dpose.py:197-207): camera-frame 3D keys carry the camera name, stacked-hourglass
# 2D detections end in "-sh"; world-frame 3D keys are "<action>.h5"
seqs = [(1, "Walking", "Walking 1.54138969.h5", 13), (1, "Eating", "Eating.55011271.h5-sh", 9),
        (5, "Sitting", "Sitting 2.60457274.h5", 21), (6, "Photo", "Photo.58860488.h5-sh", 6)]
dx = {(s, a, f): rng.standard_normal((n, 32)) for s, a, f, n in seqs}
dy_cam = {(s, a, f[:-3] if f.endswith("-sh") else f): rng.standard_normal((n, 48)) for s, a, f, n in seqs}
dy_world = {(s, a, "{0}.h5".format(f.split(".")[0])): rng.standard_normal((n, 48)) for s, a, f, n in seqs}
out["gab_n"] = np.array([n for *_, n in seqs])
out["gab_keys"] = np.array(["|".join(map(str, k[:3])) for k in seqs])
out["gab_x"] = np.vstack([dx[(s, a, f)] for s, a, f, n in seqs])
out["gab_y_cam"] = np.vstack(list(dy_cam.values()))
out["gab_y_world"] = np.vstack(list(dy_world.values()))
me = types.SimpleNamespace(input_size=32, output_size=48, batch_size=8)
for cam_frame, dy, tag in ((True, dy_cam, "cam"), (False, dy_world, "world")):
    ex, ey = ref_get_all_batches(me, dx, dy, cam_frame, training=False)
    out["gab_%s_eval_x" % tag], out["gab_%s_eval_y" % tag] = np.stack(ex), np.stack(ey)
    assert ex[0].dtype == np.float64 and len(ex) == 49 // 8
    np.random.seed(5)
    ex, ey = ref_get_all_batches(me, dx, dy, cam_frame, training=True)
    out["gab_%s_train_x" % tag], out["gab_%s_train_y" % tag] = np.stack(ex), np.stack(ey)
# a set that is an exact multiple of the batch size (n_extra == 0 branch, :291-294)
me7 = types.SimpleNamespace(input_size=32, output_size=48, batch_size=7)
ex, ey = ref_get_all_batches(me7, dx, dy_cam, True, training=False)
out["gab_cam_eval7_x"], out["gab_cam_eval7_y"] = np.stack(ex), np.stack(ey)
assert len(ex) == 7

# ------------------------------------------------------------------ evaluate_batches
src, EVB = block("/root/reference/src/predict_3dpose.py", 352, 444)
assert src[351].startswith("def evaluate_batches( sess, model,"), src[351]
assert src[443].strip() == "return total_err, joint_err, step_time, loss", src[443]

g = np.load(os.path.join(ROOT, "tests", "golden", "geometry.npz"))
t = np.load(os.path.join(ROOT, "tests", "golden", "tables.npz"))
m2, s2, use2, ig2 = g["mean2d"], g["std2d"], g["use2d"], g["ignore2d"]
m3, s3 = g["mean3d"], g["std3d"]
out.update(mean2d=m2, std2d=s2, use2d=use2, ignore2d=ig2, mean3d=m3, std3d=s3)

B, NB = 16, 3
gt96, pr96 = synth.eval_pairs(B * NB, seed=9)
pr96 = pr96.copy()
tmp = pr96[3].reshape(32, 3).copy(); tmp[:, 1] *= -1.0; pr96[3] = tmp.reshape(-1)      # a reflection case
enc = [rng.standard_normal((B, 32)) for _ in range(NB)]
out["evb_enc"] = np.stack(enc)


class StubModel:
    """Stands in for the TensorFlow session call of :396: recorded predictions (float32, as session.run returns
    them) and recorded per-batch losses."""

    def __init__(self, preds, losses):
        self.preds, self.losses, self.calls = preds, losses, 0

    def step(self, sess, enc_in, dec_out, dp, isTraining=True):
        assert dp == 1.0 and isTraining is False
        i = self.calls
        self.calls += 1
        return self.losses[i], None, self.preds[i].copy()


for p14 in (False, True):
    use3 = t["use3d_14"] if p14 else t["use3d"]
    ig3 = t["ignore3d_14"] if p14 else t["ignore3d"]
    tag = "p14" if p14 else "p17"
    gt_n = (gt96[:, use3] - m3[use3]) / s3[use3]
    pr_n = ((pr96[:, use3] - m3[use3]) / s3[use3]).astype(np.float32)
    dec = [gt_n[i * B:(i + 1) * B] for i in range(NB)]                   # float64, as get_all_batches hands them over
    preds = [pr_n[i * B:(i + 1) * B] for i in range(NB)]
    losses = [np.float32(np.mean((preds[i] - dec[i].astype(np.float32)) ** 2)) for i in range(NB)]
    out["evb_%s_dec" % tag], out["evb_%s_pred" % tag] = np.stack(dec), np.stack(preds)
    out["evb_%s_losses" % tag] = np.array(losses)
    out["evb_%s_use3d" % tag], out["evb_%s_ignore3d" % tag] = use3, ig3
    for use_proc in (False, True):
        flags = types.SimpleNamespace(batch_size=B, procrustes=use_proc, predict_14=p14)
        ns = {"np": np, "time": time, "data_utils": ref_du, "procrustes": ref_pr, "FLAGS": flags}
        exec(EVB, ns)
        model = StubModel(preds, losses)
        total_err, joint_err, step_time, loss = ns["evaluate_batches"](
            None, model, m3, s3, use3, ig3, m2, s2, use2, ig2, 0,
            [e.copy() for e in enc], [d.copy() for d in dec], current_epoch=0)
        assert model.calls == NB
        pt = "proc" if use_proc else "plain"
        out["evb_%s_%s_total" % (tag, pt)] = np.float64(total_err)
        out["evb_%s_%s_joint" % (tag, pt)] = joint_err
        out["evb_%s_%s_loss" % (tag, pt)] = np.float64(loss)
        print(tag, pt, "total_err %.6f mm, loss %.6f, joints %d" % (total_err, loss, joint_err.size))

np.savez_compressed(os.path.join(ROOT, "tests", "golden", "driver.npz"), **out)
print("wrote tests/golden/driver.npz with", len(out), "arrays,",
      os.path.getsize(os.path.join(ROOT, "tests", "golden", "driver.npz")), "bytes")
