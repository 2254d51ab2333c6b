"""The two host-side callers of the hot path against tests/golden/driver.npz, which oracle/make_golden_driver.py wrote
by EXECUTING the reference's own source lines (src/linear_model.py:247-300 get_all_batches, src/predict_3dpose.py:352-444
evaluate_batches with a stub in place of the TensorFlow session call).  CPU only: the host method and the oracle."""
import os

import numpy as np
import pytest

from oracle import geometry_ref as G


@pytest.fixture(scope="module")
def drv(golden_dir):
    return np.load(os.path.join(golden_dir, "driver.npz"))


def _dicts(drv):
    """The dictionaries the generator fed to the reference, rebuilt from the stacked arrays (same key order)."""
    keys = [tuple(k.split("|")) for k in drv["gab_keys"]]
    keys = [(int(s), a, f) for s, a, f in keys]
    ns = drv["gab_n"].tolist()
    offs = np.concatenate([[0], np.cumsum(ns)])
    dx, dy_cam, dy_world = {}, {}, {}
    for i, (s, a, f) in enumerate(keys):
        sl = slice(offs[i], offs[i + 1])
        dx[(s, a, f)] = drv["gab_x"][sl]
        dy_cam[(s, a, f[:-3] if f.endswith("-sh") else f)] = drv["gab_y_cam"][sl]
        dy_world[(s, a, "{0}.h5".format(f.split(".")[0]))] = drv["gab_y_world"][sl]
    return dx, dy_cam, dy_world


def _host_model(batch_size):
    from p3d.linear_model import LinearModel
    m = LinearModel.__new__(LinearModel)       # host-side method only: no device, no library call
    m._handle = None
    m.input_size, m.output_size, m.batch_size = 32, 48, batch_size
    return m


@pytest.mark.parametrize("tag", ["cam", "world"])
def test_get_all_batches_bit_exact(drv, tag):
    """Bit-exact (index and copy work): concatenation order, the 2D -> 3D key mapping in both coordinate frames (the
    '-sh' suffix of stacked-hourglass detections, '<action>.h5' for world-frame 3D), the permutation drawn from NumPy's
    global generator, the dropped tail, float64 batches."""
    dx, dy_cam, dy_world = _dicts(drv)
    dy, cam_frame = (dy_cam, True) if tag == "cam" else (dy_world, False)
    m = _host_model(8)
    ex, ey = m.get_all_batches(dx, dy, cam_frame, training=False)
    assert len(ex) == len(ey) == drv["gab_%s_eval_x" % tag].shape[0] == 6
    assert all(b.dtype == np.float64 and b.shape == (8, 32) for b in ex) and all(b.shape == (8, 48) for b in ey)
    assert np.array_equal(np.stack(ex), drv["gab_%s_eval_x" % tag]) and np.array_equal(np.stack(ey), drv["gab_%s_eval_y" % tag])
    np.random.seed(5)
    ex, ey = m.get_all_batches(dx, dy, cam_frame, training=True)
    assert np.array_equal(np.stack(ex), drv["gab_%s_train_x" % tag]) and np.array_equal(np.stack(ey), drv["gab_%s_train_y" % tag])


def test_get_all_batches_exact_multiple(drv):
    """49 poses in batches of 7: the n_extra == 0 branch (linear_model.py:291-294) keeps every pose."""
    dx, dy_cam, _ = _dicts(drv)
    ex, ey = _host_model(7).get_all_batches(dx, dy_cam, True, training=False)
    assert len(ex) == 7
    assert np.array_equal(np.stack(ex), drv["gab_cam_eval7_x"]) and np.array_equal(np.stack(ey), drv["gab_cam_eval7_y"])


@pytest.mark.parametrize("p14", [False, True])
@pytest.mark.parametrize("use_proc", [False, True])
def test_oracle_mpjpe_against_executed_evaluate_batches(drv, p14, use_proc):
    """The oracle's batched MPJPE (all poses at once, batched SVD) against the reference's loop (one batch at a time,
    one LAPACK SVD per pose): float64 on both sides -> 1e-9 mm."""
    tag, pt = ("p14" if p14 else "p17"), ("proc" if use_proc else "plain")
    dec = drv["evb_%s_dec" % tag].reshape(-1, 42 if p14 else 48)
    pred = drv["evb_%s_pred" % tag].reshape(dec.shape)
    assert pred.dtype == np.float32
    d = G.mpjpe(pred, dec, drv["mean3d"], drv["std3d"], drv["evb_%s_ignore3d" % tag], drv["evb_%s_use3d" % tag],
                procrustes=use_proc, predict_14=p14)
    assert d.shape == (dec.shape[0], 14 if p14 else 17)
    assert abs(d.mean() - float(drv["evb_%s_%s_total" % (tag, pt)])) < 1e-9
    np.testing.assert_allclose(d.mean(axis=0), drv["evb_%s_%s_joint" % (tag, pt)], rtol=0, atol=1e-9)
    # the reference's loss is the mean of the per-batch losses it got from model.step
    assert abs(float(np.mean(drv["evb_%s_losses" % tag])) - float(drv["evb_%s_%s_loss" % (tag, pt)])) < 1e-7
    # the oracle's index tables are the ones the reference's normalization_stats produced
    use3, ign3 = G.dims_to_use(3, p14)
    assert np.array_equal(use3, drv["evb_%s_use3d" % tag]) and np.array_equal(ign3, drv["evb_%s_ignore3d" % tag])
