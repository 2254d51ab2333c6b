"""Seeded synthetic workloads.  TEST/BENCH INFRASTRUCTURE ONLY.

There is no dataset and no network: poses, cameras and weights are synthetic, shaped like
Human3.6M (32 joint slots, 17 named, 4 cameras per subject)."""
from __future__ import annotations

import numpy as np


def mlp_inputs(B, out_size=48, seed=0, dtype=np.float32):
    rng = np.random.RandomState(seed)
    x = rng.standard_normal((B, 32)).astype(dtype)
    t = rng.standard_normal((B, out_size)).astype(dtype)
    return x, t


def random_rotation(rng):
    q, r = np.linalg.qr(rng.standard_normal((3, 3)))
    q = q * np.sign(np.diag(r))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


def cameras(ncams=4, seed=3):
    """H36M-style cameras: centre 4-6 m from the origin looking at it, f~1145 px,
    c~(512,515), k~(-0.207,0.248,-0.003), p~(-0.00098,-0.00142).  Shapes follow
    cameras.load_camera_params (cameras.py:108-115): R[3,3] T[3,1] f[2,1] c[2,1] k[3,1] p[2,1]."""
    rng = np.random.RandomState(seed)
    cams = []
    for _ in range(ncams):
        d = rng.standard_normal(3); d /= np.linalg.norm(d)
        centre = d * rng.uniform(4000.0, 6000.0)          # mm
        z = -centre / np.linalg.norm(centre)               # optical axis towards the origin
        up = np.array([0.0, 0.0, 1.0])
        if abs(z @ up) > 0.95:
            up = np.array([0.0, 1.0, 0.0])
        xax = np.cross(up, z); xax /= np.linalg.norm(xax)
        yax = np.cross(z, xax)
        R = np.stack([xax, yax, z])                         # rows = camera axes in world coords
        T = centre.reshape(3, 1)
        f = (np.array([1145.0, 1144.0]) + rng.uniform(-3, 3, 2)).reshape(2, 1)
        c = (np.array([512.0, 515.0]) + rng.uniform(-5, 5, 2)).reshape(2, 1)
        k = (np.array([-0.207, 0.248, -0.003]) * rng.uniform(0.9, 1.1, 3)).reshape(3, 1)
        p = (np.array([-0.00098, -0.00142]) * rng.uniform(0.9, 1.1, 2)).reshape(2, 1)
        cams.append((R, T, f, c, k, p))
    return cams


def world_poses(N, seed=3, dtype=np.float64):
    """[N,96]: root ~ N(0,500 mm), joints = root + N(0,300 mm); all 32 slots filled."""
    rng = np.random.RandomState(seed + 100)
    root = rng.normal(0, 500.0, (N, 1, 3))
    joints = root + rng.normal(0, 300.0, (N, 32, 3))
    joints[:, 0, :] = root[:, 0, :]
    return joints.reshape(N, 96).astype(dtype)


def eval_pairs(N, seed=4):
    """Ground truth = root-centred camera-frame-like 32-slot poses; prediction = a noisy
    similarity transform of it.  Returns gt96, pred96 float64 [N,96]."""
    rng = np.random.RandomState(seed)
    gt = rng.normal(0, 300.0, (N, 32, 3))
    gt[:, 0, :] = 0.0
    ang = rng.normal(0, 0.15, (N, 3))
    ca, sa = np.cos(ang), np.sin(ang)
    z, o = np.zeros(N), np.ones(N)
    Rx = np.stack([o, z, z, z, ca[:, 0], -sa[:, 0], z, sa[:, 0], ca[:, 0]], -1).reshape(N, 3, 3)
    Ry = np.stack([ca[:, 1], z, sa[:, 1], z, o, z, -sa[:, 1], z, ca[:, 1]], -1).reshape(N, 3, 3)
    Rz = np.stack([ca[:, 2], -sa[:, 2], z, sa[:, 2], ca[:, 2], z, z, z, o], -1).reshape(N, 3, 3)
    s = rng.uniform(0.9, 1.1, (N, 1, 1))
    pred = s * (gt @ (Rz @ Ry @ Rx)) + rng.normal(0, 20.0, (N, 1, 3)) + rng.normal(0, 30.0, (N, 32, 3))
    return gt.reshape(N, 96), pred.reshape(N, 96)
