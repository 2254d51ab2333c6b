"""GPU diagnostic for the generic tcgen05 GEMM (p3d_debug_tc_gemm, csrc/tc_gemm.cu): every operand-layout
combination the training step uses, against a float64 product of the bf16-rounded operands.  On a
mismatch it prints which 16-wide K slices / 64-wide M,N blocks are wrong (descriptor mistakes show up
as structured errors).  Optional timing with CUDA events."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-pose-baseline_b200"), os.path.join(ROOT, "tests")]
from p3d import _lib  # noqa: E402


def run(M, N, K, a_mn, b_mn, split_k=0, bias=False, res=False, colsum=False, alpha=1.0, seed=0, time_it=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn((K, M) if a_mn else (M, K), generator=g, device="cuda").to(torch.bfloat16)
    B = torch.randn((K, N) if b_mn else (N, K), generator=g, device="cuda").to(torch.bfloat16)
    C = torch.zeros((M, N), dtype=torch.float32, device="cuda") if split_k else \
        torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
    bv = torch.randn(N, generator=g, device="cuda") if bias else None
    rv = torch.randn((M, N), generator=g, device="cuda") if res else None
    cs = torch.zeros(2 * N, dtype=torch.float64, device="cuda") if colsum else None
    lda, ldb = A.shape[1], B.shape[1]

    def call():
        _lib.check(_lib.lib.p3d_debug_tc_gemm(A.data_ptr(), lda, int(a_mn), B.data_ptr(), ldb, int(b_mn), C.data_ptr(), N, M, N, K,
                                              bv.data_ptr() if bias else None, rv.data_ptr() if res else None, alpha,
                                              int(split_k), cs.data_ptr() if colsum else None, None))
    call()
    torch.cuda.synchronize()
    Af = (A.T if a_mn else A).double()
    Bf = (B.T if b_mn else B).double()
    ref = alpha * (Af @ Bf.T)
    if bias:
        ref = ref + bv.double()[None, :]
    if res:
        ref = ref + rv.double()
    got = C.double()
    scale = float(ref.abs().max())
    err = float((got - ref).abs().max()) / max(scale, 1e-30)
    msg = f"M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)} split={int(split_k)} bias={int(bias)} res={int(res)}: rel max err {err:.3e} nan={int(torch.isnan(C).sum())}"
    ok = err < 2e-5 * max(1.0, (K / 64) ** 0.5) + (1e-6 if not split_k else 1e-5)
    if colsum:
        r1, r2 = ref.sum(0), (ref * ref).sum(0)
        e1 = float((cs[:N] - r1).abs().max() / r1.abs().max().clamp_min(1e-30))
        e2 = float((cs[N:] - r2).abs().max() / r2.abs().max().clamp_min(1e-30))
        msg += f" colsum err {e1:.2e}/{e2:.2e}"
        ok = ok and e1 < 1e-4 and e2 < 1e-4
    print(msg + ("" if ok else "   <<<<<< MISMATCH"))
    if not ok:
        d = (torch.nan_to_num(got) - ref).abs()
        mb = [float(d[r:r + 64].max()) for r in range(0, min(M, 256), 64)]
        nb = [float(d[:, c:c + 64].max()) for c in range(0, min(N, 512), 64)]
        print("   err per 64-row block:", np.round(mb, 3), "\n   err per 64-col block:", np.round(nb, 3))
        ns = min(K // 16, 16)
        if ns >= 1 and M * N <= 1 << 20:
            parts = torch.stack([(Af[:, s * 16:(s + 1) * 16] @ Bf[:, s * 16:(s + 1) * 16].T).reshape(-1) for s in range(ns)], 1)
            coef = torch.linalg.lstsq(parts, torch.nan_to_num(got).reshape(-1, 1)).solution.reshape(-1)
            print("   K-slice coefficients (want all alpha):", np.round(coef.cpu().numpy(), 3))
    if time_it:
        for _ in range(3):
            call()
        e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            call()
        e1_.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1_) / 20
        print(f"   {ms * 1e3:.1f} us/launch, {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")
    return ok


if __name__ == "__main__":
    cases = [
        # layout checks on one tile
        dict(M=128, N=64, K=64, a_mn=0, b_mn=0),
        dict(M=128, N=64, K=64, a_mn=0, b_mn=1),
        dict(M=128, N=64, K=64, a_mn=1, b_mn=0),
        dict(M=128, N=64, K=64, a_mn=1, b_mn=1),
        dict(M=128, N=256, K=256, a_mn=0, b_mn=1),
        dict(M=128, N=256, K=256, a_mn=1, b_mn=1),
        # the shapes of the training step, batch 64 and 4096
        dict(M=64, N=1024, K=32, a_mn=0, b_mn=1, bias=True, colsum=True, alpha=0.37),      # z0 = x W1
        dict(M=64, N=1024, K=1024, a_mn=0, b_mn=1, bias=True, colsum=True),                # z = h W
        dict(M=64, N=48, K=1024, a_mn=0, b_mn=1, bias=True),                               # y = h W4
        dict(M=64, N=1024, K=48, a_mn=0, b_mn=0),                                          # dh = dy W4^T
        dict(M=64, N=1024, K=1024, a_mn=0, b_mn=0, res=True),                              # dh = dz W^T + dres
        dict(M=1024, N=1024, K=64, a_mn=1, b_mn=1, split_k=1),                             # dW = h^T dz
        dict(M=1024, N=48, K=64, a_mn=1, b_mn=1, split_k=1),                               # dW4
        dict(M=32, N=1024, K=64, a_mn=1, b_mn=1, split_k=1),                               # dW1
        dict(M=4096, N=1024, K=1024, a_mn=0, b_mn=1, bias=True, colsum=True, time_it=True),
        dict(M=4096, N=1024, K=1024, a_mn=0, b_mn=0, res=True, time_it=True),
        dict(M=1024, N=1024, K=4096, a_mn=1, b_mn=1, split_k=1, time_it=True),
        dict(M=4096, N=48, K=1024, a_mn=0, b_mn=1, bias=True, time_it=True),
        dict(M=2, N=1024, K=1024, a_mn=0, b_mn=0, bias=True),                              # inference layers: BN=32, 8-row A box
        dict(M=61, N=1024, K=1024, a_mn=0, b_mn=0, bias=True, res=True),
        dict(M=100, N=1024, K=64, a_mn=0, b_mn=0, bias=True),
        dict(M=1000, N=1024, K=1024, a_mn=0, b_mn=1, bias=True, colsum=True),              # ragged M
        dict(M=1024, N=1024, K=1000, a_mn=1, b_mn=1, split_k=1),                           # ragged K (batch)
        # shapes that take the CTA-pair path (cta_group::2) under P3D_GEMM_CG2=1: M >= 256 with a tile >= 128 wide
        dict(M=2176, N=1024, K=256, a_mn=0, b_mn=1, bias=True, colsum=True),               # BN=128, odd number of M tiles, MN-major B half = one box
        dict(M=2100, N=1024, K=320, a_mn=0, b_mn=0, res=True),                             # BN=128, ragged M, K-major B half = 64 rows
        dict(M=4224, N=1024, K=128, a_mn=0, b_mn=1),                                       # BN=256, 33 M tiles
        dict(M=4096, N=1024, K=48, a_mn=0, b_mn=0),                                        # dh = dy W4^T at batch 4096
        dict(M=384, N=1024, K=512, a_mn=1, b_mn=1, split_k=1),                             # split-K keeps BN=256; 3 M tiles, MN-major A
        dict(M=32768, N=1024, K=1024, a_mn=0, b_mn=1, bias=True, colsum=True, time_it=True),
        # edges of the TMA-store epilogue (P3D_GEMM_TMASTORE=1): rows and columns clipped by the tensor map of C
        dict(M=333, N=48, K=128, a_mn=0, b_mn=1, bias=True),                               # N ends inside the second 32-column box
        dict(M=97, N=1024, K=64, a_mn=0, b_mn=1, bias=True, colsum=True, alpha=0.5),       # last quadrant holds one live row
        dict(M=1000, N=96, K=1000, a_mn=1, b_mn=1, split_k=1),                             # reduce-add with ragged M, N and K
    ]
    print("P3D_GEMM_CG2 =", os.environ.get("P3D_GEMM_CG2", "0"), " P3D_GEMM_OCC2 =", os.environ.get("P3D_GEMM_OCC2", "0"),
          " P3D_GEMM_TMASTORE =", os.environ.get("P3D_GEMM_TMASTORE", "0"), " P3D_GEMM_FASTISSUE =", os.environ.get("P3D_GEMM_FASTISSUE", "0"))
    bad = 0
    for c in cases:
        try:
            bad += not run(**c)
        except Exception as e:          # a trapped kernel poisons the context: stop here
            print(f"{c}: EXCEPTION {e}")
            bad += 1
            break
    print("TCGEMM", "FAIL" if bad else "OK")
    sys.exit(1 if bad else 0)
