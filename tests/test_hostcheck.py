"""The device math header (csrc/math_hd.h) compiled for the host with g++ and compared with the
oracle/golden vectors - catches arithmetic mistakes in the kernels' inner functions without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hc():
    src = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
    out_dir = os.path.join(HERE, "hostcheck", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhostcheck.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", so, src], check=True)
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_kabsch_matches_reference_svd(hc, golden_dir):
    g = np.load(os.path.join(golden_dir, "procrustes.npz"))
    X, Y = g["X"], g["Y"]
    X0 = X - X.mean(1, keepdims=True); Y0 = Y - Y.mean(1, keepdims=True)
    nx = np.sqrt((X0 ** 2).sum((1, 2))); ny = np.sqrt((Y0 ** 2).sum((1, 2)))
    A = np.ascontiguousarray(np.swapaxes(X0 / nx[:, None, None], 1, 2) @ (Y0 / ny[:, None, None]))
    n = A.shape[0]
    T = np.zeros((n, 3, 3)); tr = np.zeros(n)
    hc.hc_kabsch(_p(A), _p(T), _p(tr), n)
    np.testing.assert_allclose(T, g["proc_T"], atol=1e-10)
    b = tr * nx / ny
    np.testing.assert_allclose(b, g["proc_b"], rtol=1e-10)
    np.testing.assert_allclose(1 - tr ** 2, g["proc_d"], atol=1e-11)
    assert np.allclose(np.linalg.det(T), 1.0)


def test_kabsch_random_and_reflections(hc):
    from oracle import geometry_ref as G
    rng = np.random.RandomState(5)
    n = 500
    X = rng.normal(0, 1, (n, 17, 3)); Y = rng.normal(0, 1, (n, 17, 3))
    Y[::3] = X[::3] * np.array([1, 1, -1.0]) + rng.normal(0, 0.05, (len(X[::3]), 17, 3))   # mirrored copies
    Y[1::7, :, 2] *= 1e-4                                                                    # nearly planar
    d, Z, Tref, bref, c = G.similarity_transform(X, Y, True)
    X0 = X - X.mean(1, keepdims=True); Y0 = Y - Y.mean(1, keepdims=True)
    nx = np.sqrt((X0 ** 2).sum((1, 2))); ny = np.sqrt((Y0 ** 2).sum((1, 2)))
    A = np.ascontiguousarray(np.swapaxes(X0 / nx[:, None, None], 1, 2) @ (Y0 / ny[:, None, None]))
    T = np.zeros((n, 3, 3)); tr = np.zeros(n)
    hc.hc_kabsch(_p(A), _p(T), _p(tr), n)
    np.testing.assert_allclose(T, Tref, atol=1e-8)
    np.testing.assert_allclose(tr * nx / ny, bref, rtol=1e-9, atol=1e-12)


def test_project_point(hc, golden_dir):
    g = np.load(os.path.join(golden_dir, "geometry.npz"))
    P = np.ascontiguousarray(g["world"].reshape(-1, 3))
    for i in range(4):
        cam = np.concatenate([g[f"cam{i}_{k}"].reshape(-1) for k in "RTfckp"]).astype(np.float64)
        out = np.zeros((P.shape[0], 6))
        hc.hc_project_f64(_p(P), _p(cam), _p(out), P.shape[0])
        np.testing.assert_allclose(out[:, :2], g[f"cam{i}_proj"], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(out[:, 2], g[f"cam{i}_D"], rtol=1e-13)
        np.testing.assert_allclose(out[:, 3], g[f"cam{i}_radial"], rtol=1e-13)
        np.testing.assert_allclose(out[:, 4], g[f"cam{i}_tan"], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(out[:, 5], g[f"cam{i}_r2"], rtol=1e-13)
        P32 = P.astype(np.float32); out32 = np.zeros((P.shape[0], 6), np.float32)
        hc.hc_project_f32(_p(P32), _p(cam), _p(out32), P.shape[0])
        # fp32 arithmetic on ~1000 px coordinates: 1e-3 px absolute
        np.testing.assert_allclose(out32[:, :2], g[f"cam{i}_proj"], atol=2e-3)
