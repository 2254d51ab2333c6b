"""Drop-in for the reference's `procrustes.compute_similarity_transform` (src/procrustes.py:2-63).

Single pose [J,3] like the reference, or a batch [N,J,3] (one CUDA thread per pose, fp64)."""
from __future__ import annotations

from . import _dev, _lib
from ._lib import lib, check


def compute_similarity_transform(X, Y, compute_optimal_scale=False):
    """Returns (d, Z, T, b, c): squared error after transformation, transformed Y, rotation, scale,
    translation (procrustes.py:2-63).  For [N,J,3] inputs every output gains a leading N."""
    torch = _lib.require_cuda()
    Xd, was = _dev.to_device(X, torch.float64)
    Yd, _ = _dev.to_device(Y, torch.float64, Xd.device)
    single = Xd.dim() == 2
    if single:
        Xd, Yd = Xd[None], Yd[None]
    if Xd.dim() != 3 or Xd.shape[2] != 3 or Xd.shape != Yd.shape:
        raise ValueError("X and Y must both be [J,3] or [N,J,3]")
    N, J = int(Xd.shape[0]), int(Xd.shape[1])
    dev = Xd.device
    d = torch.empty(N, dtype=torch.float64, device=dev)
    Z = torch.empty((N, J, 3), dtype=torch.float64, device=dev)
    T = torch.empty((N, 3, 3), dtype=torch.float64, device=dev)
    b = torch.empty(N, dtype=torch.float64, device=dev)
    c = torch.empty((N, 3), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(lib.p3d_similarity_transform_f64(Xd.data_ptr(), Yd.data_ptr(), J, int(bool(compute_optimal_scale)), N,
                                               d.data_ptr(), Z.data_ptr(), T.data_ptr(), b.data_ptr(), c.data_ptr(),
                                               _lib.current_stream()))
    outs = [d, Z, T, b, c]
    if single:
        outs = [o[0] for o in outs]
    outs = [_dev.back(o, was) for o in outs]
    if single and not was:
        outs[0] = float(outs[0]); outs[3] = float(outs[3]) if compute_optimal_scale else 1
    return tuple(outs)
