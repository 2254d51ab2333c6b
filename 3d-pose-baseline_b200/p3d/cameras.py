"""Drop-in for the math of the reference's `cameras` module (src/cameras.py:13-90) on B200.

Same signatures and return values; NumPy float64 in -> NumPy float64 out (computed by fp64 CUDA
kernels), or torch CUDA tensors in -> torch out without host copies.  The HDF5 loaders
(load_camera_params / load_cameras, cameras.py:92-138) are dataset I/O and out of scope."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _dev, _lib
from ._lib import lib, check


def _points(P):
    if len(P.shape) != 2 or P.shape[1] != 3:         # cameras.py:35-36
        raise AssertionError("P must be Nx3")


def project_point_radial(P, R, T, f, c, k, p):
    """Project points from 3d to 2d using camera parameters including radial and tangential
    distortion (cameras.py:13-53).  Returns (Proj[N,2], D[N], radial[N], tan[N], r2[N])."""
    _points(P)
    torch = _lib.require_cuda()
    Pd, was = _dev.to_device(P, torch.float64)
    n = Pd.shape[0]
    cam = _lib.make_camera(R, T, f, c, k, p)
    proj = torch.empty((n, 2), dtype=torch.float64, device=Pd.device)
    aux = torch.empty((4, n), dtype=torch.float64, device=Pd.device)
    with torch.cuda.device(Pd.device):
        check(lib.p3d_project_point_radial_f64(Pd.data_ptr(), C.byref(cam), proj.data_ptr(), aux[0].data_ptr(),
                                               aux[1].data_ptr(), aux[2].data_ptr(), aux[3].data_ptr(), n,
                                               _lib.current_stream()))
    return tuple(_dev.back(t, was) for t in (proj, aux[0], aux[1], aux[2], aux[3]))


def _rigid(fn, P, R, T):
    _points(P)
    torch = _lib.require_cuda()
    Pd, was = _dev.to_device(P, torch.float64)
    cam = _lib.make_camera(R, T, [1.0, 1.0], [0.0, 0.0], [0.0, 0.0, 0.0], [0.0, 0.0])
    out = torch.empty_like(Pd)
    with torch.cuda.device(Pd.device):
        check(fn(Pd.data_ptr(), C.byref(cam), out.data_ptr(), Pd.shape[0], _lib.current_stream()))
    return _dev.back(out, was)


def world_to_camera_frame(P, R, T):
    """Convert points from world to camera coordinates (cameras.py:55-72)."""
    return _rigid(lib.p3d_world_to_camera_f64, P, R, T)


def camera_to_world_frame(P, R, T):
    """Inverse of world_to_camera_frame (cameras.py:74-90)."""
    return _rigid(lib.p3d_camera_to_world_f64, P, R, T)
