"""Parity of the CUDA training step (model.step(isTraining=True), src/linear_model.py:129-145,203-235)
with the fp64 oracle: forward with batch statistics + dropout, loss, gradients through clip_by_norm /
BN / ReLU / dropout / residual, BN moving averages, TF-Adam + lr decay.

The kernels compute in fp32: after a few steps the variables agree with the fp64 oracle to ~1e-4
relative to each tensor's scale (Adam's m/sqrt(v) normalisation amplifies fp32 rounding of tiny
gradients, so biases that feed a BatchNorm - whose true gradient is exactly 0, SURVEY appendix A.4 -
are pinned to "do not move at all")."""
import os

import numpy as np
import pytest
import torch

from oracle import mlp_ref as M
from oracle import synth
from helpers import make_model

pytestmark = pytest.mark.gpu


def _close(a, b, rtol, name):
    scale = max(np.abs(b).max(), 1e-6)
    err = np.abs(np.asarray(a, np.float64) - b).max()
    assert err <= rtol * scale, (name, err, scale)


CFGS = [M.Config(1024, 2, True, True, True), M.Config(256, 2, True, True, False), M.Config(128, 1, False, True, True),
        M.Config(64, 2, True, False, False), M.Config(128, 0, True, True, True)]


@pytest.mark.parametrize("cfg", CFGS, ids=lambda c: f"L{c.linear_size}n{c.num_layers}r{int(c.residual)}b{int(c.batch_norm)}m{int(c.max_norm)}")
def test_train_steps_match_oracle_with_injected_masks(cfg):
    B, keep, lr0, steps = 64, 0.5, 1e-3, 3
    m, p = make_model(cfg, seed=21, bn="fresh", mode="fp32", lr=lr0)
    st = M.AdamState()
    nh = 2 * cfg.num_layers + 1
    rng = np.random.RandomState(0)
    for s in range(steps):
        x, t = synth.mlp_inputs(B, seed=50 + s)
        masks = (rng.uniform(size=(nh, B, cfg.linear_size)) < keep).astype(np.uint8)
        loss, _, lr_sum, y = m.step(None, x, t, keep, isTraining=True, dropout_mask=masks)
        rloss, rlr, ry = M.train_step(p, st, x.astype(np.float64), t.astype(np.float64), cfg, lr0, keep_prob=keep,
                                      masks=list(masks))
        assert abs(float(loss) - rloss) <= 2e-4 * max(1.0, rloss), (s, float(loss), rloss)
        assert abs(float(lr_sum.value) - rlr) <= 1e-6 * rlr
        _close(y, ry, 2e-4, f"outputs step {s}")
    assert int(m.global_step) == steps
    got = m.get_variables(include_optimizer=True)
    for name in M.trainable_names(cfg.linear_size, cfg.num_layers, cfg.batch_norm):
        leaf = name.rsplit("/", 1)[-1]
        if cfg.batch_norm and leaf[0] == "b" and leaf != "b4" and "batch_normalization" not in name:
            # bias feeding a BN layer: zero gradient, Adam must leave it bit-identical
            assert np.array_equal(got[name], p[name].astype(np.float32)), name
            continue
        # Every element moved by ~lr per step.  Adam's m/(sqrt(v)+eps) turns fp32 rounding of near-zero
        # gradients into O(lr) differences on a few elements, so: 99.5% of the elements within 2% of
        # the total movement, and nothing further off than the movement itself.
        err = np.abs(got[name].astype(np.float64) - p[name])
        move = lr0 * steps
        assert np.quantile(err, 0.995) <= 0.02 * move + 1e-6 * np.abs(p[name]).max(), (name, np.quantile(err, 0.995))
        assert err.max() <= 2.2 * move, (name, err.max())
    if cfg.batch_norm:
        for name in p:
            if name.endswith(("moving_mean", "moving_variance")):
                _close(got[name], p[name], 1e-4, name)
    m.close()


@pytest.mark.parametrize("cfg", CFGS, ids=lambda c: f"L{c.linear_size}n{c.num_layers}r{int(c.residual)}b{int(c.batch_norm)}m{int(c.max_norm)}")
def test_gradients_match_oracle(cfg):
    """model.gradients (linear_model.py:143-144) after one step == the oracle's backward pass at the
    pre-update variables: 2e-4 of each tensor's largest gradient."""
    B, keep = 64, 0.5
    m, p = make_model(cfg, seed=31, bn="trained", mode="fp32", lr=1e-3)
    nh = 2 * cfg.num_layers + 1
    x, t = synth.mlp_inputs(B, seed=77)
    masks = (np.random.RandomState(4).uniform(size=(nh, B, cfg.linear_size)) < keep).astype(np.uint8)
    m.step(None, x, t, keep, isTraining=True, dropout_mask=masks)
    got = m.get_gradients()
    y, cache = M.forward(p, x.astype(np.float64), cfg, training=True, keep_prob=keep, masks=list(masks), want_cache=True)
    grads = M.backward(p, x.astype(np.float64), t.astype(np.float64), cfg, cache, y)
    for name, g in grads.items():
        scale = np.abs(g).max()
        if scale < 1e-12:                 # biases in front of a BN layer: exactly zero
            assert np.abs(got[name]).max() == 0.0, name
            continue
        assert np.abs(got[name].astype(np.float64) - g).max() <= 2e-4 * scale, (name, np.abs(got[name] - g).max(), scale)
    m.close()


TC_CFGS = [M.Config(1024, 2, True, True, True), M.Config(1024, 2, True, True, False), M.Config(256, 1, False, True, True),
           M.Config(64, 2, True, False, False), M.Config(128, 0, True, True, True)]


@pytest.mark.parametrize("fused", ["1", "0"], ids=["fused-epilogues", "unfused"])
@pytest.mark.parametrize("B", [64, 200, 1024])
@pytest.mark.parametrize("cfg", TC_CFGS, ids=lambda c: f"L{c.linear_size}n{c.num_layers}r{int(c.residual)}b{int(c.batch_norm)}m{int(c.max_norm)}")
def test_bf16_tensor_core_step_matches_oracle(cfg, B, fused, monkeypatch):
    """mode='bf16': every GEMM of the step runs on tcgen05 with bf16 operands (fp32 accumulate, fp32 master
    weights).
    (1) Against the oracle restated with the SAME rounding points (oracle forward/backward(quant=bf16)).  Without
        BatchNorm the kernel reproduces it to fp32 accuracy (measured 1e-7..3e-4; bound 1e-3 relative L2, 1% of the
        largest entry).  With BatchNorm the fp32-vs-fp64 difference of the normalised activations (1e-6) decides the
        bf16 rounding direction of ~5e-4 of the activations and a few ReLU derivatives, so the match is statistical:
        outputs within 3e-3, gradients within 5e-2 relative L2 (measured 4e-4..2.7e-2).
    (2) Against the exact fp64 graph: loss and outputs within 1e-2 (north_star's bf16 tolerance).  Gradients are
        only loosely comparable there: operand rounding moves pre-activations by ~2^-9, which flips the ReLU
        derivative of ~0.3% of the units, i.e. ~sqrt(0.003) = 5% in relative L2 (measured 3-10%) - so: <= 20%.
    Biases in front of a BatchNorm stay EXACTLY zero-gradient.
    Both routes: GEMMs with the BatchNorm / ReLU / dropout arithmetic fused into their epilogues (one M tile at B = 64,
    grid-synchronised at 200 and 1024) and, with P3D_TRAIN_FUSED=0, the unfused route that larger batches take."""
    if B > 64 and cfg.linear_size > 256 and not cfg.max_norm:
        pytest.skip("covered at B=64")
    if fused == "0" and (B == 64 or cfg.linear_size == 64):
        pytest.skip("unfused route covered at B=200/1024 on the wider models")
    monkeypatch.setenv("P3D_TRAIN_FUSED", fused)
    from helpers import bf16_round
    keep = 0.5
    m, p = make_model(cfg, seed=31, bn="trained", mode="bf16", lr=1e-3)
    nh = 2 * cfg.num_layers + 1
    x, t = synth.mlp_inputs(B, seed=77)
    x64, t64 = x.astype(np.float64), t.astype(np.float64)
    masks = (np.random.RandomState(4).uniform(size=(nh, B, cfg.linear_size)) < keep).astype(np.uint8)
    loss, _, _, yk = m.step(None, x, t, keep, isTraining=True, dropout_mask=masks)
    got = m.get_gradients()
    m.close()
    # (2) exact graph
    y, cache = M.forward(p, x64, cfg, training=True, keep_prob=keep, masks=list(masks), want_cache=True)
    rloss = M.loss_fn(y, t64)
    assert abs(float(loss) - rloss) <= 1e-2 * max(1.0, rloss), (float(loss), rloss)
    assert np.abs(yk - y).max() <= 1e-2 * max(np.abs(y).max(), 1.0)
    grads = M.backward(p, x64, t64, cfg, cache, y)
    # (1) same rounding points
    q = lambda a: bf16_round(a).astype(np.float64)  # noqa: E731
    yq, cq = M.forward(p, x64, cfg, training=True, keep_prob=keep, masks=list(masks), want_cache=True, quant=q)
    gq = M.backward(p, x64, t64, cfg, cq, yq, quant=q)
    assert abs(float(loss) - M.loss_fn(yq, t64)) <= 1e-4 * max(1.0, rloss)
    assert np.abs(yk - yq).max() <= (3e-3 if cfg.batch_norm else 1e-4) * max(np.abs(yq).max(), 1.0)
    for name, g in grads.items():
        scale = np.abs(g).max()
        if scale < 1e-12:
            assert np.abs(got[name]).max() == 0.0, name
            continue
        k = got[name].astype(np.float64)
        assert np.linalg.norm(k - g) <= 0.2 * np.linalg.norm(g), (name, "exact graph", np.linalg.norm(k - g) / np.linalg.norm(g))
        d = k - gq[name]
        tol = 5e-2 if cfg.batch_norm else 1e-3
        assert np.linalg.norm(d) <= tol * np.linalg.norm(gq[name]), (name, "same rounding", np.linalg.norm(d) / np.linalg.norm(gq[name]))
        if not cfg.batch_norm:
            assert np.abs(d).max() <= 1e-2 * np.abs(gq[name]).max(), (name, "same rounding", np.abs(d).max(), np.abs(gq[name]).max())


def test_bf16_and_fp32_training_trajectories_agree():
    """20 Adam steps from the same initial state in both modes: the loss curves stay within 2%."""
    cfg = M.Config(1024, 2, True, True, True)
    losses = {}
    for mode in ("fp32", "bf16"):
        m, _ = make_model(cfg, seed=3, bn="fresh", mode=mode, lr=1e-3)
        out = []
        for s in range(20):
            x, t = synth.mlp_inputs(256, seed=100 + s)
            masks = (np.random.RandomState(s).uniform(size=(5, 256, 1024)) < 0.5).astype(np.uint8)
            loss, _, _, _ = m.step(None, x, t, 0.5, isTraining=True, dropout_mask=masks)
            out.append(float(loss))
        losses[mode] = np.array(out)
        m.close()
    assert np.all(np.isfinite(losses["bf16"]))
    assert np.abs(losses["bf16"] - losses["fp32"]).max() <= 2e-2 * losses["fp32"].max(), (losses["bf16"], losses["fp32"])
    assert losses["fp32"][-1] < losses["fp32"][0]


def test_gradients_via_single_step_sgd_like_probe():
    """One step from zero Adam state moves every variable by -alpha*sign(g) (|m|/sqrt(v) = 1 up to eps):
    the sign pattern of the update must equal the sign of the oracle gradient wherever |g| is not tiny."""
    cfg = M.Config(256, 1, True, True, True)
    m, p = make_model(cfg, seed=5, bn="fresh", mode="fp32", lr=1e-3)
    x, t = synth.mlp_inputs(64, seed=3)
    before = m.get_variables()
    m.step(None, x, t, 1.0, isTraining=True)
    after = m.get_variables()
    y, cache = M.forward(p, x.astype(np.float64), cfg, training=True, want_cache=True)
    grads = M.backward(p, x.astype(np.float64), t.astype(np.float64), cfg, cache, y)
    for name in ("linear_model/w1", "linear_model/two_linear_0/w2_0", "linear_model/w4", "linear_model/b4",
                 "linear_model/batch_normalization/gamma", "linear_model/two_linear_0/batch_normalization20/beta"):
        g = grads[name]
        big = np.abs(g) > 1e-3 * np.abs(g).max()
        delta = after[name].astype(np.float64) - before[name]
        assert np.array_equal(np.sign(delta[big]), -np.sign(g[big])), name
        alpha = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
        expect = -alpha * 0.1 * g / (np.sqrt(0.001 * g * g) + 1e-8)         # TF-Adam, first step, eps included
        assert np.allclose(delta[big], expect[big], rtol=2e-2, atol=2e-6), name
    m.close()


def test_generated_dropout_mask_is_the_documented_philox_stream():
    """Without an injected mask the kernel draws floor(keep + u) from Philox4x32-10 keyed by
    (seed; global_row, col/4, layer, global_step): the oracle generates the same mask, so a full
    training step must agree; and the step is reproducible for a fixed seed."""
    cfg = M.Config(128, 1, True, True, False)
    B, keep = 32, 0.5
    x, t = synth.mlp_inputs(B, seed=8)
    outs = []
    for _ in range(2):
        m, p = make_model(cfg, seed=77, bn="fresh", mode="fp32", lr=1e-3)
        loss, _, _, y = m.step(None, x, t, keep, isTraining=True)
        outs.append((float(loss), y))
        seed = m._seed
        m.close()
    assert outs[0][0] == outs[1][0] and np.array_equal(outs[0][1], outs[1][1])
    masks = [M.dropout_mask_philox(seed, 0, li, B, cfg.linear_size, keep) for li in range(3)]
    st = M.AdamState()
    rloss, _, ry = M.train_step(p, st, x.astype(np.float64), t.astype(np.float64), cfg, 1e-3, keep_prob=keep, masks=masks)
    assert abs(outs[0][0] - rloss) <= 2e-4 * max(1.0, rloss)
    _close(outs[0][1], ry, 2e-4, "outputs")
    frac = np.mean([mk.mean() for mk in masks])
    assert 0.45 < frac < 0.55


def test_training_reduces_loss_and_inference_uses_moving_stats():
    cfg = M.Config(1024, 2, True, True, True)
    m, _ = make_model(cfg, seed=9, bn="fresh", mode="bf16", lr=1e-3)
    rng = np.random.RandomState(0)
    W = rng.standard_normal((32, 48)).astype(np.float32) * 0.3
    x = rng.standard_normal((4096, 32)).astype(np.float32)
    t = x @ W
    xd, td = torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda()
    first = None
    for it in range(60):
        idx = torch.randint(0, 4096, (256,), device="cuda")
        loss, _, _, _ = m.step(None, xd[idx], td[idx], 0.9, isTraining=True)
        first = first if first is not None else float(loss)
    assert float(loss) < 0.5 * first
    # eval path works right after training (weights re-folded with the updated moving statistics)
    # (after 60 steps the moving statistics are still 55% their initial values (momentum 0.99), so the inference
    #  loss is not yet below the first training loss: require sanity, and that the statistics moved)
    l_eval, _, y = m.step(None, xd, td, 1.0, isTraining=False)
    assert torch.isfinite(y).all() and float(l_eval) < 2.0 * first
    mm = m.get_variables()["linear_model/batch_normalization/moving_mean"]
    assert np.abs(mm).max() > 1e-3
    m.close()


def test_checkpoint_roundtrip(tmp_path):
    cfg = M.Config(128, 1, True, True, True)
    m, _ = make_model(cfg, seed=1, bn="fresh", mode="fp32")
    x, t = synth.mlp_inputs(64, seed=1)
    m.step(None, x, t, 0.8, isTraining=True)
    path = str(tmp_path / "checkpoint-1.npz")
    m.save(path)
    _, _, y_a = m.step(None, x, t, 1.0, isTraining=False)
    m2, _ = make_model(cfg, seed=99, bn="fresh", mode="fp32")
    m2.restore(path)
    _, _, y_b = m2.step(None, x, t, 1.0, isTraining=False)
    assert np.array_equal(y_a, y_b) and int(m2.global_step) == 1
    with pytest.raises(ValueError):
        m2.restore(str(tmp_path / "checkpoint-2.npz"))
    m.close(); m2.close()


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_train_epoch_equals_stepping_through_the_batches(mode):
    """p3d_model_train_epoch (device-side gather + one replayed CUDA graph per batch) == the reference's loop
    `for i in range(nbatches): model.step(sess, enc[i], dec[i], keep, isTraining=True)` (predict_3dpose.py:242-249)
    over the same permutation; the ragged tail is dropped like get_all_batches does (linear_model.py:311-313)."""
    cfg = M.Config(256, 1, True, True, True)
    n, B, keep = 1000, 64, 0.7
    x, t = synth.mlp_inputs(n, seed=12)
    perm = np.random.RandomState(3).permutation(n)
    ma, _ = make_model(cfg, seed=5, bn="fresh", mode=mode, batch_size=B)
    mb, _ = make_model(cfg, seed=5, bn="fresh", mode=mode, batch_size=B)
    losses, lr = ma.train_epoch(x, t, keep, perm=perm)
    assert losses.shape == (n // B,) and int(ma.global_step) == n // B
    ref = []
    for b in range(n // B):
        idx = perm[b * B:(b + 1) * B]
        loss, _, lrs, _ = mb.step(None, x[idx], t[idx], keep, isTraining=True)
        ref.append(float(loss))
    np.testing.assert_allclose(losses, np.array(ref), rtol=2e-4, atol=1e-6)
    assert abs(lr - float(lrs.value)) < 1e-9
    va, vb = ma.get_variables(), mb.get_variables()
    for k in va:
        assert np.abs(va[k] - vb[k]).max() <= 2e-3 * max(np.abs(vb[k]).max(), 1e-3), k    # Adam amplifies last-bit differences of atomics
    # torch in -> torch out, natural order
    lt, _ = ma.train_epoch(torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda(), 1.0, shuffle=False)
    assert lt.is_cuda and lt.shape == (n // B,) and torch.isfinite(lt).all()
    ma.close(); mb.close()


@pytest.mark.parametrize("mode,B", [("bf16", 64), ("bf16", 200), ("fp32", 64)])
def test_predict_14_training_step(mode, B):
    """predict_14 (output width 42, linear_model.py:69): the output layer's bf16 operands are padded to a 48-wide
    pitch for TMA; loss, outputs and the w4/b4 gradients must still match the oracle (fused route at B=64, unfused at 200)."""
    from helpers import bf16_round
    cfg = M.Config(256, 1, True, True, True, out_size=42)
    m, p = make_model(cfg, seed=17, bn="trained", mode=mode, predict_14=True)
    x, _ = synth.mlp_inputs(B, seed=5)
    t = np.random.RandomState(1).standard_normal((B, 42)).astype(np.float32)
    masks = (np.random.RandomState(2).uniform(size=(3, B, 256)) < 0.8).astype(np.uint8)
    loss, _, _, y = m.step(None, x, t, 0.8, isTraining=True, dropout_mask=masks)
    got = m.get_gradients()
    m.close()
    q = (lambda a: bf16_round(a).astype(np.float64)) if mode == "bf16" else None
    x64, t64 = x.astype(np.float64), t.astype(np.float64)
    yr, cache = M.forward(p, x64, cfg, training=True, keep_prob=0.8, masks=list(masks), want_cache=True, quant=q)
    gr = M.backward(p, x64, t64, cfg, cache, yr, quant=q)
    tol_y, tol_g = (3e-3, 5e-2) if mode == "bf16" else (2e-4, 1e-3)
    assert y.shape == (B, 42)
    assert abs(float(loss) - M.loss_fn(yr, t64)) <= 1e-3 * max(1.0, float(loss))
    assert np.abs(y - yr).max() <= tol_y * max(np.abs(yr).max(), 1.0)
    for name in ("linear_model/w4", "linear_model/b4", "linear_model/w1", "linear_model/two_linear_0/w3_0"):
        d = got[name].astype(np.float64) - gr[name]
        assert np.linalg.norm(d) <= tol_g * np.linalg.norm(gr[name]), (name, np.linalg.norm(d) / np.linalg.norm(gr[name]))
