"""GPU diagnostic: where does the bf16 kernel differ from its NumPy emulation?  Prints error quantiles
(relative to the output rms) for several depths and emulation variants."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-pose-baseline_b200"), os.path.join(ROOT, "tests")]
from helpers import bf16_round, make_model  # noqa: E402
from oracle import mlp_ref as M, synth  # noqa: E402


def trunc_bf16(a):
    a32 = np.ascontiguousarray(np.asarray(a, np.float32))
    return (a32.view(np.uint32) & 0xFFFF0000).view(np.float32).reshape(a32.shape)


def emulate(p, x, cfg, act_round=bf16_round, w_round=bf16_round, acc=np.float64):
    folded = M.fold_inference(p, cfg)
    h = act_round(x).astype(acc)
    res = None
    last = len(folded) - 1
    hs = []
    for li, (w, b) in enumerate(folded):
        w = w_round(w.astype(np.float32)).astype(acc)
        z = (h @ w).astype(np.float64) + b.astype(np.float32).astype(np.float64)
        if li == last:
            return z, hs
        r = np.maximum(z, 0)
        r = act_round(r).astype(acc)
        h = r if (li == 0 or li % 2 == 1) else ((res + r) if cfg.residual else r)
        h = act_round(h).astype(acc)
        hs.append(h)
        if li == 0 or li % 2 == 0:
            res = h


def q(err, rms):
    return " ".join(f"{np.quantile(err, v) / rms:.2e}" for v in (0.5, 0.9, 0.99, 1.0))


for maxnorm in (False, True):
    for nl in (0, 1, 2):
        cfg = M.Config(1024, nl, True, True, maxnorm)
        m, p = make_model(cfg, seed=11, bn="trained", mode="bf16")
        x, t = synth.mlp_inputs(128, seed=128)
        _, _, y = m.step(None, x, t, 1.0, isTraining=False)
        ref = M.forward(p, x.astype(np.float64), cfg, training=False)
        rms = np.sqrt(np.mean(ref ** 2))
        e_rn, _ = emulate(p, x, cfg)
        e_tr, _ = emulate(p, x, cfg, act_round=trunc_bf16)
        e_f32, _ = emulate(p, x, cfg, acc=np.float32)
        print(f"max_norm={maxnorm} nl={nl} rms={rms:.3f}  quantiles(50,90,99,100) of |y-emu|/rms:")
        print("   vs RN emulation   :", q(np.abs(y - e_rn), rms))
        print("   vs trunc-act emu  :", q(np.abs(y - e_tr), rms))
        print("   vs fp32-acc emu   :", q(np.abs(y - e_f32), rms))
        print("   RN emu vs oracle  :", q(np.abs(e_rn - ref), rms), "  kernel vs oracle:", q(np.abs(y - ref), rms))
        rb = [np.quantile(np.abs(y - e_rn)[r:r + 32], 0.9) / rms for r in range(0, 128, 32)]
        print("   p90 err per 32-row block:", " ".join(f"{v:.2e}" for v in rb))
        # fp32-mode kernel as a third opinion
        m32, _ = make_model(cfg, seed=11, bn="trained", mode="fp32")
        _, _, y32 = m32.step(None, x, t, 1.0, isTraining=False)
        print("   fp32 kernel vs oracle:", q(np.abs(y32 - ref), rms))
        m.close(); m32.close()
