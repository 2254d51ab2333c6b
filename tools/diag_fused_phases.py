"""In-kernel phase times (globaltimer stamps per CTA) of the fused training GEMMs of one step:
python tools/diag_fused_phases.py [B] [layer]   - forward (mode 3) and backward (mode 4) kernel of hidden layer `layer`."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-pose-baseline_b200")]
from p3d import LinearModel
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
layer = int(sys.argv[2]) if len(sys.argv) > 2 else 1
os.environ["P3D_TRAIN_GRAPH"] = "0"            # plans are made per step: the environment switches below take effect
names = {0: "entry", 1: "setup done", 2: "producer issued all", 3: "first tile landed", 4: "all MMAs issued", 14: "phase 0 done (under mainloop)",
         5: "accumulator complete", 15: "pass 1 done", 16: "column totals known", 6: "epilogue done", 7: "block end"}
for mode in (3, 4):
    m = LinearModel(1024, 2, True, True, True, B, 1e-3, seed=1, mode="bf16")
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((B, 32), device="cuda", generator=g); t = torch.randn((B, 48), device="cuda", generator=g)
    for _ in range(3):
        m.step(None, x, t, 0.5, isTraining=True)
    dbg = torch.zeros((4096, 32), dtype=torch.int64, device="cuda")
    os.environ.update(P3D_GEMM_DBG_PTR=str(dbg.data_ptr()), P3D_GEMM_DBG_MODE=str(mode), P3D_GEMM_DBG_LAYER=str(layer))
    m.step(None, x, t, 0.5, isTraining=True)
    torch.cuda.synchronize()
    for k in ("P3D_GEMM_DBG_PTR", "P3D_GEMM_DBG_MODE", "P3D_GEMM_DBG_LAYER"):
        del os.environ[k]
    d = dbg.cpu().numpy()
    d = d[d[:, 0] > 0]
    t0 = d[:, 0].min()
    rel = d - t0
    print(f"B={B} hidden layer {layer} {'forward (mode 3)' if mode == 3 else 'backward (mode 4)'}: {len(d)} CTAs; kernel span {(d[:, 7].max() - t0) / 1e3:.1f} us")
    for i in (0, 1, 3, 2, 4, 14, 5, 15, 16, 6, 7):
        if d[:, i].max() <= 0:
            continue
        print(f"   {names[i]:32s} mean {rel[:, i].mean() / 1e3:6.2f} us   min {rel[:, i].min() / 1e3:6.2f}   max {rel[:, i].max() / 1e3:6.2f}")
    m.close()
