"""The arithmetic of the reference's `predict_3dpose.evaluate_batches` (src/predict_3dpose.py:399-442)
as one fused CUDA pass: un-normalise ground truth and prediction, keep [0,1,2] U dim_to_use_3d
(17 joints incl. the hip), optional per-pose Procrustes alignment (procrustes.py:2-63 with
compute_optimal_scale=True, re-applied as b*out.dot(T)+c), per-joint Euclidean error."""
from __future__ import annotations

import time

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


def evaluate_batches(sess, model,
                     data_mean_3d, data_std_3d, dim_to_use_3d, dim_to_ignore_3d,
                     data_mean_2d, data_std_2d, dim_to_use_2d, dim_to_ignore_2d,
                     current_step, encoder_inputs, decoder_outputs, current_epoch=0, *, procrustes=False, dist=None):
    """`predict_3dpose.evaluate_batches` (src/predict_3dpose.py:352-444) with its signature: evaluates a list of
    batches (what `model.get_all_batches(..., training=False)` returns) and gives back
    (total_err, joint_err, step_time, loss).  `--procrustes` is a keyword here (the reference reads the global FLAGS,
    :413), `predict_14` is taken from the model.

    The reference walks the batches one `session.run` at a time with a Python loop per pose inside; here the batches
    are stacked once, uploaded once, lifted by one `model.step` and scored by one `mpjpe` pass - every pose is
    independent at test time (moving BatchNorm statistics, keep probability 1), so the numbers are the same.  Like
    the reference (:410-411,433) every batch must hold `model.batch_size` poses; `loss` is the mean of the per-batch
    losses, which for equal batches is the mean over all elements."""
    nbatches = len(encoder_inputs)
    if nbatches == 0 or len(decoder_outputs) != nbatches:
        raise ValueError("encoder_inputs / decoder_outputs must be equally long, non-empty lists of batches")
    for e, d in zip(encoder_inputs, decoder_outputs):
        assert e.shape[0] == model.batch_size and d.shape[0] == model.batch_size, "every batch must hold model.batch_size poses"
    torch = _lib.require_cuda()
    start_time = time.time()
    dev = torch.device("cuda", model.device)
    if hasattr(encoder_inputs[0], "is_cuda"):
        enc, dec = torch.cat(list(encoder_inputs), 0), torch.cat(list(decoder_outputs), 0)
    else:
        enc = torch.from_numpy(np.concatenate([np.asarray(e, dtype=np.float32) for e in encoder_inputs], 0)).to(dev)
        dec = torch.from_numpy(np.concatenate([np.asarray(d, dtype=np.float32) for d in decoder_outputs], 0)).to(dev)
    loss, _, poses3d = model.step(sess, enc, dec, 1.0, isTraining=False)      # dropout keep probability is 1 at test time
    total_err, joint_err = mpjpe(poses3d, dec, data_mean_3d, data_std_3d, procrustes=procrustes,
                                 predict_14=model.predict_14, dist=dist)
    loss = float(loss)
    step_time = (time.time() - start_time) / nbatches
    return total_err, joint_err, step_time, loss
