"""Drop-in for the data math of the reference's `data_utils` module on B200
(src/data_utils.py:195-311,339-364,474-494): normalisation statistics, normalise / un-normalise,
root-centring and the per-camera projection / camera-frame drivers.

Same signatures, same dictionary keys, NumPy float64 results - computed by libp3d CUDA kernels.
The HDF5/CDF dataset readers (load_data, load_stacked_hourglass, read_3d_data, create_2d_data,
read_2d_predictions: data_utils.py:61-192,367-471) are dataset I/O and out of scope.

`camera_frame_dataset` is the fused fp32 fast path of the same pipeline (one pass over the world
poses for all cameras)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _dev, _lib, cameras
from ._lib import lib, check

# Human3.6m IDs for training and testing (data_utils.py:15-16)
TRAIN_SUBJECTS = [1, 5, 6, 7, 8]
TEST_SUBJECTS = [9, 11]

# Joints in H3.6M -- data has 32 joints, but only 17 that move (data_utils.py:19-39)
H36M_NAMES = [""] * 32
for _i, _n in {0: "Hip", 1: "RHip", 2: "RKnee", 3: "RFoot", 6: "LHip", 7: "LKnee", 8: "LFoot", 12: "Spine",
               13: "Thorax", 14: "Neck/Nose", 15: "Head", 17: "LShoulder", 18: "LElbow", 19: "LWrist",
               25: "RShoulder", 26: "RElbow", 27: "RWrist"}.items():
    H36M_NAMES[_i] = _n

# Stacked Hourglass produces 16 joints (data_utils.py:42-59)
SH_NAMES = ["RFoot", "RKnee", "RHip", "LHip", "LKnee", "LFoot", "Hip", "Spine", "Thorax", "Head", "RWrist",
            "RElbow", "RShoulder", "LShoulder", "LElbow", "LWrist"]


def _dims(dim, predict_14=False):
    named = np.where(np.array([x != "" for x in H36M_NAMES]))[0]
    if dim == 2:
        j = np.array([i for i in named if H36M_NAMES[i] != "Neck/Nose"])
        use = np.sort(np.hstack((j * 2, j * 2 + 1)))
        total = len(H36M_NAMES) * 2
    elif dim == 3:
        j = np.delete(named, [0, 7, 9] if predict_14 else 0)
        use = np.sort(np.hstack((j * 3, j * 3 + 1, j * 3 + 2)))
        total = len(H36M_NAMES) * 3
    else:
        raise ValueError("dim must be 2 or 3")
    return use, np.delete(np.arange(total), use)


def normalization_stats(complete_data, dim, predict_14=False):
    """Mean, population stdev, dimensions ignored / used (data_utils.py:195-230)."""
    use, ignore = _dims(dim, predict_14)
    torch = _lib.require_cuda()
    d, _ = _dev.to_device(complete_data, torch.float64)
    n, D = int(d.shape[0]), int(d.shape[1])
    out = torch.empty((2, D), dtype=torch.float64, device=d.device)
    work = torch.empty(2 * D, dtype=torch.float64, device=d.device)
    with torch.cuda.device(d.device):
        check(lib.p3d_column_stats_f64(d.data_ptr(), n, D, out[0].data_ptr(), out[1].data_ptr(), work.data_ptr(),
                                       _lib.current_stream()))
    o = out.cpu().numpy()
    return o[0].copy(), o[1].copy(), ignore, use


def _mode_from_use(dim_to_use, full_width=None):
    """The kernels carry the reference's three index tables; find which one the caller passed."""
    u = np.asarray(dim_to_use).reshape(-1)
    for dim, p14 in ((2, False), (3, False), (3, True)):
        if u.size == _dims(dim, p14)[0].size and np.array_equal(u, _dims(dim, p14)[0]):
            return dim, p14
    raise ValueError("dim_to_use is not one of the H36M tables produced by normalization_stats")


def _normalize_array(arr, data_mean, data_std, dim, p14):
    torch = _lib.require_cuda()
    d, was = _dev.to_device(arr, torch.float64)
    n = int(d.shape[0])
    out = torch.empty((n, len(_dims(dim, p14)[0])), dtype=torch.float64, device=d.device)
    m, s = _dev.host_f64(data_mean), _dev.host_f64(data_std)
    with torch.cuda.device(d.device):
        check(lib.p3d_normalize_f64(d.data_ptr(), _lib.np_ptr(m), _lib.np_ptr(s), dim, int(p14), out.data_ptr(), n,
                                    _lib.current_stream()))
    return _dev.back(out, was)


def normalize_data(data, data_mean, data_std, dim_to_use):
    """Normalizes a dictionary of poses (data_utils.py:260-280).  Like the reference it also replaces
    data[key] by its dim_to_use columns (:275)."""
    dim, p14 = _mode_from_use(dim_to_use)
    data_out = {}
    for key in data.keys():
        full = data[key]
        data_out[key] = _normalize_array(full, data_mean, data_std, dim, p14)
        data[key] = full[:, dim_to_use]
    return data_out


def unNormalizeData(normalized_data, data_mean, data_std, dimensions_to_ignore):
    """Un-normalizes a matrix; ignored dimensions come back as the mean (data_utils.py:283-311)."""
    D = int(np.asarray(data_mean).shape[0])
    ign = np.asarray(dimensions_to_ignore).reshape(-1)
    dim = 2 if D == 64 else 3
    p14 = None
    for cand in ((False, True) if dim == 3 else (False,)):
        if np.array_equal(np.sort(ign), _dims(dim, cand)[1]):
            p14 = cand
    if D not in (64, 96) or p14 is None:
        raise ValueError("dimensions_to_ignore is not one of the H36M tables produced by normalization_stats")
    torch = _lib.require_cuda()
    d, was = _dev.to_device(normalized_data, torch.float64)
    n = int(d.shape[0])
    if d.shape[1] != D - ign.size:
        raise ValueError("normalized_data has %d columns, expected %d" % (d.shape[1], D - ign.size))
    out = torch.empty((n, D), dtype=torch.float64, device=d.device)
    m, s = _dev.host_f64(data_mean), _dev.host_f64(data_std)
    with torch.cuda.device(d.device):
        check(lib.p3d_unnormalize_f64(d.data_ptr(), _lib.np_ptr(m), _lib.np_ptr(s), dim, int(p14), out.data_ptr(), n,
                                      _lib.current_stream()))
    return _dev.back(out, was)


def postprocess_3d(poses_set):
    """Center 3d points around root (data_utils.py:474-494).  Returns (poses_set, root_positions)."""
    torch = _lib.require_cuda()
    root_positions = {}
    for k in poses_set.keys():
        d, was = _dev.to_device(poses_set[k], torch.float64)
        n, D = int(d.shape[0]), int(d.shape[1])
        out = torch.empty_like(d)
        roots = torch.empty((n, 3), dtype=torch.float64, device=d.device)
        with torch.cuda.device(d.device):
            check(lib.p3d_root_center_f64(d.data_ptr(), out.data_ptr(), roots.data_ptr(), n, D, _lib.current_stream()))
        root_positions[k] = _dev.back(roots, was)
        poses_set[k] = _dev.back(out, was)
    return poses_set, root_positions


def _per_subject_all_cameras(poses_set, cams, ncams, kernel, out_cols):
    """Shared driver of project_to_cameras / transform_world_to_camera: the reference loops over sequences x cameras
    (data_utils.py:243-255, 349-362) with one NumPy call each.  Here all sequences of a subject (they share their
    cameras) travel to the device as ONE fp64 buffer, every camera is one kernel launch over all of its points, and
    the ncams results come back in ONE copy - two PCIe transfers per subject instead of 2 x sequences x cameras.
    `kernel(P_dev_ptr, camera, out_dev_ptr, npts, stream)`; out_cols = values per point (2 or 3)."""
    torch = _lib.require_cuda()
    out = {}
    keys = sorted(poses_set.keys())
    by_subject = {}
    for k in keys:
        by_subject.setdefault(k[0], []).append(k)
    nj = len(H36M_NAMES)
    for subj, ks in by_subject.items():
        is_torch = hasattr(poses_set[ks[0]], "is_cuda")
        rows = [int(poses_set[k].shape[0]) for k in ks]
        if is_torch:
            P = torch.cat([poses_set[k].reshape(-1, 3).to(torch.float64) for k in ks], 0).contiguous().cuda()
        else:
            P = torch.from_numpy(np.ascontiguousarray(np.concatenate([np.reshape(np.asarray(poses_set[k], dtype=np.float64), [-1, 3])
                                                                      for k in ks], 0))).cuda()
        npts = int(P.shape[0])
        res = torch.empty((ncams, npts, out_cols), dtype=torch.float64, device=P.device)
        names = []
        with torch.cuda.device(P.device):
            for cam in range(ncams):
                R, T, f, c, k_, p_, name = cams[(subj, cam + 1)]
                names.append(name)
                kernel(P.data_ptr(), _lib.make_camera(R, T, f, c, k_, p_), res[cam].data_ptr(), npts, _lib.current_stream())
        res_h = res if is_torch else res.cpu().numpy()
        start = 0
        for k, n in zip(ks, rows):
            _, a, seqname = k
            for cam in range(ncams):
                blk = res_h[cam, start * nj:(start + n) * nj]
                out[(subj, a, seqname[:-3] + "." + names[cam] + ".h5")] = blk.reshape(n, nj * out_cols)
            start += n
    return out


def project_to_cameras(poses_set, cams, ncams=4):
    """Project 3d poses using camera parameters (data_utils.py:339-364)."""
    def kernel(P, cam, out, npts, stream):
        check(lib.p3d_project_point_radial_f64(P, C.byref(cam), out, None, None, None, None, npts, stream))
    return _per_subject_all_cameras(poses_set, cams, ncams, kernel, 2)


def transform_world_to_camera(poses_set, cams, ncams=4):
    """Project 3d poses from world coordinate to camera coordinate system (data_utils.py:233-257)."""
    def kernel(P, cam, out, npts, stream):
        check(lib.p3d_world_to_camera_f64(P, C.byref(cam), out, npts, stream))
    return _per_subject_all_cameras(poses_set, cams, ncams, kernel, 3)


def camera_frame_dataset(world, cams, data_mean_2d=None, data_std_2d=None, data_mean_3d=None, data_std_3d=None,
                         predict_14=False, want_2d=True, want_3d=True):
    """Fused fp32 preprocessing: world poses [N,96] x cameras -> normalised network inputs
    x2d[ncams,N,32] (project_to_cameras + normalize_data) and targets y3d[ncams,N,48|42]
    (transform_world_to_camera + postprocess_3d + normalize_data) in ONE pass over `world`.
    `cams`: list of (R,T,f,c,k,p[,name]).  NumPy in -> NumPy out, torch CUDA in -> torch out."""
    torch = _lib.require_cuda()
    w, was = _dev.to_device(world, torch.float32)
    N = int(w.shape[0])
    if w.dim() != 2 or w.shape[1] != 96:
        raise ValueError("world must be [N,96]")
    ncams = len(cams)
    carr = (_lib.Camera * ncams)(*[_lib.make_camera(*cam[:6]) for cam in cams])
    x2d = torch.empty((ncams, N, 32), dtype=torch.float32, device=w.device) if want_2d else None
    y3d = torch.empty((ncams, N, 42 if predict_14 else 48), dtype=torch.float32, device=w.device) if want_3d else None
    m2 = _dev.host_f64(data_mean_2d) if want_2d else None
    s2 = _dev.host_f64(data_std_2d) if want_2d else None
    m3 = _dev.host_f64(data_mean_3d) if want_3d else None
    s3 = _dev.host_f64(data_std_3d) if want_3d else None
    with torch.cuda.device(w.device):
        check(lib.p3d_project_normalize(w.data_ptr(), carr, ncams,
                                        _lib.np_ptr(m2) if want_2d else None, _lib.np_ptr(s2) if want_2d else None,
                                        _lib.np_ptr(m3) if want_3d else None, _lib.np_ptr(s3) if want_3d else None,
                                        int(predict_14), x2d.data_ptr() if want_2d else None,
                                        y3d.data_ptr() if want_3d else None, N, _lib.current_stream()))
    return (_dev.back(x2d, was) if want_2d else None), (_dev.back(y3d, was) if want_3d else None)
