"""Realtime front-end: the oracle restatement, the host-side list handling of p3d.realtime and the
device arithmetic header (compiled for the host) against golden vectors produced by EXECUTING the reference's own
source lines (oracle/make_golden_realtime.py).  Everything here is bit-exact: it is fp64 index work and IEEE ops."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import geometry_ref as G, realtime_ref as R

CASES = ["coco54", "tfpose36", "body25_75", "wide_87"]
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "realtime.npz"))


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_lines(gold, case):
    m2, s2, use2, ig2 = gold["mean2d"], gold["std2d"], gold["use2d"], gold["ignore2d"]
    m3, s3, ig3 = gold["mean3d"], gold["std3d"], gold["ignore3d"]
    enc_prev = None
    for f in range(3):
        xy = R.keypoints_to_xy(gold["%s_kp%d" % (case, f)].tolist())
        assert np.array_equal(np.array(xy, dtype=np.float64), gold["%s_xy%d" % (case, f)])
        enc, sx, sy = R.frontend(xy, m2, s2, use2, enc_prev)
        assert np.array_equal(enc, gold["%s_enc%d" % (case, f)])
        assert np.array_equal(np.array([sx, sy]), gold["%s_spine%d" % (case, f)])
        pose = G.unnormalize(gold["%s_y%d" % (case, f)], m3, s3, ig3)
        assert np.array_equal(pose, gold["%s_pose%d" % (case, f)])
        assert np.array_equal(R.display_transform(pose, sx, sy), gold["%s_display%d" % (case, f)])
        enc_prev = G.unnormalize(enc, m2, s2, ig2)      # what the reference carries into the next frame (:170)


def test_carry_over_never_reaches_the_network_input(gold):
    """The reference re-uses enc_in across frames; every used 2D dimension is overwritten each frame, so the
    stateless device front-end (zeros for unwritten joints) gives the same network input."""
    m2, s2, use2 = gold["mean2d"], gold["std2d"], gold["use2d"]
    xy = gold["coco54_xy1"].tolist()
    a, _, _ = R.frontend(xy, m2, s2, use2, None)
    b, _, _ = R.frontend(xy, m2, s2, use2, np.random.RandomState(0).normal(0, 500, (1, 64)))
    assert np.array_equal(a, b)
    written = {j for j in R.ORDER} | {0, 13, 14}
    assert {int(d) // 2 for d in use2} <= written


@pytest.mark.parametrize("case", CASES)
def test_host_list_handling_of_the_product(gold, case):
    from p3d import realtime as P
    assert P.ORDER == gold["order"].tolist()
    for f in range(3):
        xy = P.keypoints_to_xy(gold["%s_kp%d" % (case, f)].tolist())
        assert np.array_equal(np.array(xy, dtype=np.float64), gold["%s_xy%d" % (case, f)])
        sx, sy = gold["%s_spine%d" % (case, f)]
        assert xy[2] == sx and xy[3] == sy              # Spine = OpenPose keypoint 1 (:157-158)
        np.testing.assert_array_equal(P.display_transform(gold["%s_pose%d" % (case, f)], sx, sy), gold["%s_display%d" % (case, f)])


def test_short_keypoint_list_raises():
    from p3d import realtime as P
    assert len(P.keypoints_to_xy(list(range(30)))) == 30     # passes through; indexing 36 entries fails downstream


def test_device_header_front_end_is_bit_exact(gold):
    src = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
    out_dir = os.path.join(HERE, "hostcheck", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhostcheck_rt.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", so, src], check=True)
    hc = ctypes.CDLL(so)
    use2 = np.ascontiguousarray(gold["use2d"], dtype=np.int32)
    m2, s2 = np.ascontiguousarray(gold["mean2d"]), np.ascontiguousarray(gold["std2d"])
    kp = np.ascontiguousarray(np.stack([gold["%s_xy%d" % (c, f)][:36] for c in CASES for f in range(3)]))
    want = np.concatenate([gold["%s_enc%d" % (c, f)] for c in CASES for f in range(3)])
    enc = np.zeros((kp.shape[0], 32))
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    hc.hc_openpose_frontend(p(kp), p(use2), p(m2), p(s2), p(enc), kp.shape[0])
    assert np.array_equal(enc, want)
    e64 = np.zeros(64)
    hc.hc_openpose_h36m64(p(kp[0]), p(e64))
    ref = np.zeros((1, 64))
    for i, j in enumerate(R.ORDER):
        ref[0, j * 2:j * 2 + 2] = kp[0, i * 2:i * 2 + 2]
    ref[0, 0:2] = (ref[0, 2:4] + ref[0, 12:14]) / 2
    ref[0, 28:30] = (ref[0, 30:32] + ref[0, 24:26]) / 2
    ref[0, 26:28] = 2 * ref[0, 24:26] - ref[0, 28:30]
    assert np.array_equal(e64, ref[0])
