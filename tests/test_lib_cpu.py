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


@pytest.mark.parametrize("threads", [1, 3, 0])
def test_host_pack_bf16_is_the_device_rounding(threads):
    """p3d_host_pack_bf16 (host threads, no GPU): bit-identical to round-to-nearest-even fp32 -> bf16 as torch / the
    device's cvt.rn.bf16.f32 do it - ties, subnormals, overflow to inf, signed zeros; NaN becomes the canonical 0x7FFF."""
    import ctypes as C

    import torch
    from p3d import _lib
    rng = np.random.RandomState(3)
    n = 1 << 18 if threads != 1 else 70001
    x = rng.standard_normal(n).astype(np.float32) * np.float32(10.0) ** rng.randint(-30, 30, n).astype(np.float32)
    special = np.array([0.0, -0.0, np.inf, -np.inf, 3.3895314e38, -3.4e38, 1e-40, -1e-45, 1.0, 1.00390625, 1.01171875,
                        1.0078125, -1.00390625], dtype=np.float32)       # 1 + 2^-8, 1 + 3*2^-8: exact ties (even / odd neighbour)
    bits = rng.randint(0, 1 << 32, 4096, dtype=np.uint64).astype(np.uint32)
    x[:special.size] = special
    x[special.size:special.size + bits.size] = bits.view(np.float32)       # arbitrary bit patterns, NaNs among them
    out = np.empty(n, dtype=np.uint16)
    _lib.check(_lib.lib.p3d_host_pack_bf16(x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), n, threads))
    ref = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    nan = np.isnan(x)
    assert np.array_equal(out[~nan], ref[~nan])
    assert nan.any() and np.all(out[nan] == 0x7FFF)
    _lib.check(_lib.lib.p3d_host_pack_bf16(None, None, 0, threads))         # empty input


def test_host_pack_bf16_from_concurrent_callers():
    """The library's host thread pool runs one job at a time; callers from several threads queue up behind each other
    (distinct model handles may be driven from distinct threads, include/p3d.h) - no deadlock, no torn output."""
    import ctypes as C
    import threading

    import torch
    from p3d import _lib
    rng = np.random.RandomState(5)
    xs = [rng.standard_normal(300000 + 1000 * i).astype(np.float32) for i in range(4)]
    outs = [np.zeros(x.size, dtype=np.uint16) for x in xs]

    def work(i):
        for _ in range(10):
            _lib.check(_lib.lib.p3d_host_pack_bf16(xs[i].ctypes.data_as(C.c_void_p), outs[i].ctypes.data_as(C.c_void_p), xs[i].size, 0))

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in th)
    for x, o in zip(xs, outs):
        assert np.array_equal(o, torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16))
