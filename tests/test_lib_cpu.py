"""CPU-side checks: the C-ABI library loads and exports every symbol include/p3d.h declares, fails
loudly without a GPU (no CPU fallback), and the host-side logic of the Python mirror."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "p3d.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(p3d_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from p3d import _lib
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(_lib.lib, s), f"libp3d.so does not export {s}"
        assert s in _lib.PROTOTYPES, f"{s} has no ctypes prototype"
    assert _lib.lib.p3d_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from p3d import LinearModel, _lib, cameras
    with pytest.raises(_lib.P3DError):
        LinearModel(1024, 2, True, True, True, 64, 1e-3)
    with pytest.raises(_lib.P3DError):
        cameras.project_point_radial(np.zeros((4, 3)), np.eye(3), np.zeros((3, 1)), np.ones((2, 1)), np.zeros((2, 1)),
                                     np.zeros((3, 1)), np.zeros((2, 1)))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "3d-pose-baseline_b200", "p3d")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn
            assert "/root/reference" not in src, fn


def test_shard_rows_partition():
    from p3d.linear_model import shard_rows
    for n in (0, 1, 7, 64, 1000, 1 << 20):
        for w in (1, 2, 3, 4, 8):
            cuts = [shard_rows(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_kaiming_matches_reference_statistics():
    from p3d.linear_model import kaiming
    w = kaiming((1024, 1024), np.random.RandomState(0))
    assert np.abs(w).max() <= 2.0 * np.sqrt(2 / 1024) + 1e-7
    # truncated normal at 2 sigma has std 0.8796: ||W||_F ~ 39.8 for a 1024x1024 matrix
    assert abs(w.std() / np.sqrt(2 / 1024) - 0.8796) < 5e-3
    assert abs(np.linalg.norm(w) - 39.8) < 0.3


def test_dims_tables_match_golden(golden_dir):
    from p3d import data_utils
    t = np.load(os.path.join(golden_dir, "tables.npz"))
    for dim, p14, u, i in [(2, False, "use2d", "ignore2d"), (3, False, "use3d", "ignore3d"),
                           (3, True, "use3d_14", "ignore3d_14")]:
        use, ign = data_utils._dims(dim, p14)
        assert np.array_equal(use, t[u]) and np.array_equal(ign, t[i])
    perm = [data_utils.SH_NAMES.index(h) for h in data_utils.H36M_NAMES if h != "" and h in data_utils.SH_NAMES]
    assert perm == t["sh_to_gt_perm"].tolist()


def test_get_all_batches_contract():
    """linear_model.py:247-300: concatenation order, key mapping, tail dropped, fp64, fixed-size batches."""
    from p3d.linear_model import LinearModel
    m = LinearModel.__new__(LinearModel)       # host-side method only: no device needed
    m._handle = None
    m.input_size, m.output_size, m.batch_size = 32, 48, 8
    rng = np.random.RandomState(0)
    dx = {(1, "Walking", "Walking 1.54138969.h5"): rng.rand(13, 32), (1, "Eating", "Eating.55011271.h5-sh"): rng.rand(9, 32)}
    dy = {(1, "Walking", "Walking 1.54138969.h5"): rng.rand(13, 48), (1, "Eating", "Eating.55011271.h5"): rng.rand(9, 48)}
    ex, ey = m.get_all_batches(dx, dy, camera_frame=True, training=False)
    assert len(ex) == 2 and all(b.shape == (8, 32) for b in ex) and all(b.shape == (8, 48) for b in ey)
    allx = np.vstack(list(dx.values()))[:16]
    assert np.array_equal(np.vstack(ex), allx) and ex[0].dtype == np.float64
    np.random.seed(0)
    ex2, ey2 = m.get_all_batches(dx, dy, camera_frame=True, training=True)
    np.random.seed(0)
    perm = np.random.permutation(22)
    assert np.array_equal(np.vstack(ex2), np.vstack(list(dx.values()))[perm][:16])
