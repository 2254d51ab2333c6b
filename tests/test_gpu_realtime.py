"""GPU parity of the realtime front-end (src/openpose_3dpose_sandbox_realtime.py:137-171): the fused
one-launch frame step and the batched device step against golden vectors made by executing the reference's own lines
(front-end, un-normalisation) and against the MLP oracle (the lifter in between)."""
import os

import numpy as np
import pytest

from helpers import emulate_bf16_forward, make_model, rowwise_rel
from oracle import geometry_ref as G, mlp_ref as M, realtime_ref as R

pytestmark = pytest.mark.gpu
CASES = ["coco54", "tfpose36", "body25_75", "wide_87"]


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "realtime.npz"))


def _lifter(model, gold, use3=None):
    from p3d.realtime import RealtimeLifter
    return RealtimeLifter(model, gold["mean2d"], gold["std2d"], gold["use2d"], gold["mean3d"], gold["std3d"],
                          gold["use3d"] if use3 is None else use3)


@pytest.mark.parametrize("width,mode", [(1024, "bf16"), (256, "bf16"), (1024, "fp32")])
def test_frame_step_matches_reference_lines(gold, width, mode):
    """width 1024 / bf16 = the single fused cluster-kernel launch; the other two take the staged route."""
    from p3d.realtime import keypoints_to_xy
    cfg = M.Config(width, 2, True, True, True)
    model, p = make_model(cfg, seed=3, mode=mode)
    lifter = _lifter(model, gold)
    m3, s3, ig3 = gold["mean3d"], gold["std3d"], gold["ignore3d"]
    for case in CASES:
        for f in range(3):
            xy = keypoints_to_xy(gold["%s_kp%d" % (case, f)].tolist())
            enc, y, pose = lifter.step(xy)
            want_enc = gold["%s_enc%d" % (case, f)]
            # front-end: fp64 IEEE arithmetic then the feed's fp32 cast - bit exact
            assert np.array_equal(enc, want_enc.astype(np.float32))
            # back-end: unNormalizeData of the network's own fp32 output - bit exact
            assert np.array_equal(pose, G.unnormalize(y, m3, s3, ig3))
            # lifter in between: the oracle forward on the same normalised input
            ref = M.forward(p, enc.astype(np.float64), cfg, training=False)
            if mode == "fp32":
                assert rowwise_rel(y, ref).max() <= 1e-4
            else:
                assert rowwise_rel(y, ref).max() <= 1e-2          # north-star bf16 tolerance
                emu = emulate_bf16_forward(p, enc, cfg, small_batch=(width == 1024))
                assert rowwise_rel(y, emu).max() <= 2e-3
            # < 0.5 mm after un-normalisation (north star) against the oracle's un-normalised prediction
            assert np.abs(pose - G.unnormalize(ref.astype(np.float32), m3, s3, ig3)).max() < (0.5 if mode == "bf16" else 0.05)
    lifter.close()
    model.close()


def test_frame_step_equals_plain_step(gold):
    """The fused launch computes exactly what model.step computes on the same normalised input."""
    cfg = M.Config(1024, 2, True, True, True)
    model, _ = make_model(cfg, seed=4)
    lifter = _lifter(model, gold)
    xy = gold["coco54_xy0"].tolist()
    enc, y, pose = lifter.step(xy)
    _, _, y2 = model.step(None, enc, np.zeros((1, 48)), 1.0, isTraining=False)
    assert np.array_equal(y, y2)
    for _ in range(50):                                  # the flag protocol over many frames
        e2, y3, p3 = lifter.step(xy)
        assert np.array_equal(y3, y) and np.array_equal(p3, pose) and np.array_equal(e2, enc)
    # parameters changed -> the next frame sees the refreshed fold
    model.set_variable("linear_model/b4", np.zeros(48, dtype=np.float32))
    _, y4, _ = lifter.step(xy)
    _, _, y5 = model.step(None, enc, np.zeros((1, 48)), 1.0, isTraining=False)
    assert np.array_equal(y4, y5) and not np.array_equal(y4, y)
    lifter.close()
    model.close()


@pytest.mark.parametrize("B", [1, 7, 300])
def test_batched_frames(gold, B):
    cfg = M.Config(1024, 2, True, True, True)
    model, p = make_model(cfg, seed=5)
    lifter = _lifter(model, gold)
    rng = np.random.RandomState(B)
    kp = rng.uniform(50, 950, size=(B, 36))
    enc, y, pose = lifter.step_batch(kp)
    want = np.concatenate([R.frontend(kp[i].tolist(), gold["mean2d"], gold["std2d"], gold["use2d"])[0] for i in range(B)])
    assert np.array_equal(enc, want.astype(np.float32))
    assert np.array_equal(pose, G.unnormalize(y, gold["mean3d"], gold["std3d"], gold["ignore3d"]))
    ref = M.forward(p, enc.astype(np.float64), cfg, training=False)
    assert rowwise_rel(y, ref).max() <= 1e-2
    lifter.close()
    model.close()


def test_predict_14_and_bad_tables(gold, golden_dir):
    from p3d import _lib
    tab = np.load(os.path.join(golden_dir, "tables.npz"))
    cfg = M.Config(1024, 2, True, True, True)
    model, _ = make_model(cfg, seed=6, predict_14=True)
    lifter = _lifter(model, gold, use3=tab["use3d_14"])
    enc, y, pose = lifter.step(gold["tfpose36_xy0"].tolist())
    assert y.shape == (1, 42)
    assert np.array_equal(pose, G.unnormalize(y, gold["mean3d"], gold["std3d"], tab["ignore3d_14"]))
    with pytest.raises(IndexError):
        lifter.step(list(range(30)))
    lifter.close()
    with pytest.raises(ValueError):
        _lifter(model, gold)                             # 48 used dims for a 42-wide model
    bad = gold["use3d"].copy()[:42]
    bad[1] = bad[0]
    with pytest.raises(_lib.P3DError):
        _lifter(model, gold, use3=bad)                   # repeated dimension
    model.close()


def test_posebase_facade_matches_linear_model():
    """src/top_vae_3d_pose/models.py:287-481 - the TF2 twin's call signature over the same kernels."""
    from p3d.linear_model import PoseBase
    cfg = M.Config(1024, 2, True, True, True)
    p = {k: v.astype(np.float32) for k, v in M.init_params(1024, 2, seed=8, bn="trained").items()}
    pb = PoseBase(units=1024, input_size=32, output_size=48, seed=1)
    pb.load_weights(p)
    x = np.random.RandomState(0).normal(size=(33, 32))
    y = pb(x, training=False)
    ref = M.forward({k: v.astype(np.float64) for k, v in p.items()}, x.astype(np.float32).astype(np.float64), cfg, training=False)
    assert y.shape == (33, 48) and rowwise_rel(y, ref).max() <= 1e-2
    assert np.array_equal(pb.w2_1, p["linear_model/two_linear_1/w2_1"]) and pb.b4.shape == (48,)
    with pytest.raises(NotImplementedError):
        pb(x, training=True)
    with pytest.raises(AttributeError):
        pb.w5
    pb.close()
