"""GPU diagnostic: per-tensor gradient error of one training step in bf16 (tcgen05) and fp32 (FFMA) mode
against the fp64 oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-pose-baseline_b200"), os.path.join(ROOT, "tests")]
from helpers import make_model  # noqa: E402
from oracle import mlp_ref as M  # noqa: E402
from oracle import synth  # noqa: E402


def run(cfg, B, keep=0.5):
    nh = 2 * cfg.num_layers + 1
    x, t = synth.mlp_inputs(B, seed=77)
    masks = (np.random.RandomState(4).uniform(size=(nh, B, cfg.linear_size)) < keep).astype(np.uint8)
    res = {}
    for mode in ("fp32", "bf16"):
        m, p = make_model(cfg, seed=31, bn="trained", mode=mode, lr=1e-3)
        loss, _, _, yk = m.step(None, x, t, keep, isTraining=True, dropout_mask=masks)
        res[mode] = (float(loss), yk, m.get_gradients())
        m.close()
    y, cache = M.forward(p, x.astype(np.float64), cfg, training=True, keep_prob=keep, masks=list(masks), want_cache=True)
    grads = M.backward(p, x.astype(np.float64), t.astype(np.float64), cfg, cache, y)
    from helpers import bf16_round
    q = lambda a: bf16_round(a).astype(np.float64)  # noqa: E731
    yq, cq = M.forward(p, x.astype(np.float64), cfg, training=True, keep_prob=keep, masks=list(masks), want_cache=True, quant=q)
    gq = M.backward(p, x.astype(np.float64), t.astype(np.float64), cfg, cq, yq, quant=q)
    print(f"   same-rounding oracle: loss {np.mean((yq - t) ** 2):.6f} |y_bf16 - yq|max {np.abs(res['bf16'][1] - yq).max():.2e}")
    print(f"== L={cfg.linear_size} nl={cfg.num_layers} res={cfg.residual} bn={cfg.batch_norm} clip={cfg.max_norm} B={B}: "
          f"loss oracle {np.mean((y - t) ** 2):.6f} fp32 {res['fp32'][0]:.6f} bf16 {res['bf16'][0]:.6f}; "
          f"|y-yref|max fp32 {np.abs(res['fp32'][1] - y).max():.2e} bf16 {np.abs(res['bf16'][1] - y).max():.2e} (|y|max {np.abs(y).max():.2f})")
    for name, g in grads.items():
        sc = np.abs(g).max()
        if sc < 1e-12:
            continue
        row = f"  {name:58s} |g|max {sc:.2e}"
        for mode in ("fp32", "bf16"):
            d = res[mode][2][name].astype(np.float64) - g
            row += f"  {mode}: max {np.abs(d).max() / sc:.2e} relL2 {np.linalg.norm(d) / np.linalg.norm(g):.2e}"
        d = res["bf16"][2][name].astype(np.float64) - gq[name]
        row += f"  bf16 vs same-rounding: max {np.abs(d).max() / np.abs(gq[name]).max():.2e} relL2 {np.linalg.norm(d) / np.linalg.norm(gq[name]):.2e}"
        print(row)


if __name__ == "__main__":
    run(M.Config(1024, 2, True, True, True), 64)
    run(M.Config(1024, 2, True, True, True), 1024)
    run(M.Config(1024, 2, True, True, False), 64)
    run(M.Config(64, 2, True, False, False), 64)
