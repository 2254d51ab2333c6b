"""Device-memory plumbing shared by the geometry/evaluation mirrors (torch is used for allocation,
copies and streams only - every arithmetic op is a libp3d kernel)."""
from __future__ import annotations

import numpy as np

from . import _lib


def to_device(a, dtype, device=None):
    """NumPy array or torch tensor -> contiguous CUDA tensor of `dtype` (a torch dtype).
    Returns (tensor, was_torch)."""
    torch = _lib.require_cuda()
    if hasattr(a, "is_cuda"):
        t = a if a.is_cuda else a.cuda(device)
        return t.to(dtype).contiguous(), True
    npdt = {torch.float64: np.float64, torch.float32: np.float32}[dtype]
    t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=npdt)))
    return t.cuda(device), False


def back(t, was_torch):
    return t if was_torch else t.cpu().numpy()


def host_f64(a):
    return np.ascontiguousarray(np.asarray(a.cpu() if hasattr(a, "is_cuda") else a, dtype=np.float64).reshape(-1))
