"""The arithmetic of the reference's `predict_3dpose.evaluate_batches` (src/predict_3dpose.py:399-442)
as one fused CUDA pass: un-normalise ground truth and prediction, keep [0,1,2] U dim_to_use_3d
(17 joints incl. the hip), optional per-pose Procrustes alignment (procrustes.py:2-63 with
compute_optimal_scale=True, re-applied as b*out.dot(T)+c), per-joint Euclidean error."""
from __future__ import annotations

import numpy as np

from . import _dev, _lib
from ._lib import lib, check


def mpjpe(poses3d_n, dec_out_n, data_mean_3d, data_std_3d, procrustes=False, predict_14=False,
          return_dists=False, dist=None, precision="fp32"):
    """poses3d_n / dec_out_n: normalised prediction / ground truth [N,48|42] (NumPy or torch CUDA).
    Returns (total_err, joint_err[J]) in mm, plus dists[N,J] if return_dists.
    `dist`: a torch.distributed module/process group owner - when given, the per-joint sums and the
    pose count are all-reduced so that every rank returns the global error.
    `precision`: "fp32" = the HBM-bound kernel (means within 1e-5 mm of the float64 reference, single
    distances within 1e-3 mm); "fp64" = all alignment arithmetic in double (distances within 1e-6 mm)."""
    if precision not in ("fp32", "fp64"):
        raise ValueError("precision must be 'fp32' or 'fp64'")
    kernel = lib.p3d_procrustes_mpjpe if precision == "fp32" else lib.p3d_procrustes_mpjpe_f64
    torch = _lib.require_cuda()
    pd, _ = _dev.to_device(poses3d_n, torch.float32)
    gd, _ = _dev.to_device(dec_out_n, torch.float32, pd.device)
    width = 42 if predict_14 else 48
    if pd.dim() != 2 or pd.shape[1] != width or pd.shape != gd.shape:
        raise ValueError("expected two [N,%d] arrays" % width)
    N = int(pd.shape[0])
    J = 14 if predict_14 else 17
    mean = _dev.host_f64(data_mean_3d); std = _dev.host_f64(data_std_3d)
    if mean.size != 96 or std.size != 96:
        raise ValueError("data_mean_3d / data_std_3d must have 96 entries")
    sums = torch.zeros(J + 1, dtype=torch.float64, device=pd.device)
    dists = torch.empty((N, J), dtype=torch.float32, device=pd.device) if return_dists else None
    with torch.cuda.device(pd.device):
        check(kernel(pd.data_ptr(), gd.data_ptr(), _lib.np_ptr(mean), _lib.np_ptr(std),
                                       int(predict_14), int(bool(procrustes)), N,
                                       dists.data_ptr() if return_dists else None, sums.data_ptr(),
                                       _lib.current_stream()))
    sums[J] = N
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        red = sums if dist.get_backend() == "nccl" else sums.cpu()
        dist.all_reduce(red)
        sums = red
    s = sums.cpu().numpy()
    n_all = s[J]
    joint_err = s[:J] / n_all                   # np.mean(all_dists, axis=0)   (predict_3dpose.py:441)
    total_err = s[:J].sum() / (n_all * J)       # np.mean(all_dists)           (:442)
    if return_dists:
        return total_err, joint_err, dists
    return total_err, joint_err
