"""In-kernel phase times of tc_gemm (globaltimer stamps per CTA): where do the ~25 us of a 4096x1024x1024 GEMM go?"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-pose-baseline_b200")]
from p3d import _lib
def run(M, N, K, a_mn, b_mn, split_k=0, colsum=False, res=False):
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda").to(torch.bfloat16)
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda").to(torch.bfloat16)
    C = torch.zeros((M, N), device="cuda")
    cs = torch.zeros(2 * N, dtype=torch.float64, device="cuda") if colsum else None
    rv = torch.randn((M, N), device="cuda") if res else None
    dbg = torch.zeros((4096, 32), dtype=torch.int64, device="cuda")
    def call():
        _lib.check(_lib.lib.p3d_debug_tc_gemm(A.data_ptr(), A.shape[1], a_mn, B.data_ptr(), B.shape[1], b_mn, C.data_ptr(), N, M, N, K,
                                              None, rv.data_ptr() if res else None, 1.0, split_k, cs.data_ptr() if colsum else None, None))
    for _ in range(3): call()
    os.environ["P3D_GEMM_DBG_PTR"] = str(dbg.data_ptr())
    call(); torch.cuda.synchronize()
    del os.environ["P3D_GEMM_DBG_PTR"]
    d = dbg.cpu().numpy()
    d = d[d[:, 0] > 0]
    t0 = d[:, 0].min()
    rel = d - t0
    names = ["entry", "setup done", "producer issued all", "first tile landed", "all MMAs issued", "accumulator complete", "epilogue done", "block end",
             "producer issued kb 3", "producer issued kb 7", "producer issued kb 11", "tile 4 landed+MMA free", "tile 8 landed+MMA free", "tile 12 landed+MMA free"]
    print(f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} split={split_k} colsum={colsum} res={res}: {len(d)} CTAs; kernel span {(d[:,7].max()-t0)/1e3:.1f} us")
    for i, n in enumerate(names):
        if rel[:, i].max() < 0: continue
        print(f"   {n:22s} mean {rel[:, i].mean()/1e3:6.2f} us   min {rel[:, i].min()/1e3:6.2f}   max {rel[:, i].max()/1e3:6.2f}")
run(4096, 1024, 1024, 0, 1, colsum=True)
run(4096, 1024, 1024, 0, 0, res=True)
run(1024, 1024, 4096, 1, 1, split_k=1)
run(64, 1024, 1024, 0, 1, colsum=True)
run(64, 1024, 64, 0, 1)
