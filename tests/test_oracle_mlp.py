"""The MLP oracle (oracle/mlp_ref.py) against an independent torch-autograd statement of the
same TF graph (fp64).  The reference's TensorFlow code cannot run here (parity unpinned at the
TF boundary); this checks the hand-written backward, the clip_by_norm gradient, batch-stat BN,
TF-Adam and the inference folding for self-consistency."""
import numpy as np
import pytest
import torch

from oracle import mlp_ref as M


def torch_forward(tp, x, cfg, training, masks=None, keep=1.0):
    names = M.layer_names(cfg.num_layers)
    h = x
    res = None
    for li, (wn, bn, bns) in enumerate(names):
        w = tp[wn]
        if cfg.max_norm:
            nrm = torch.sqrt((w * w).sum())
            w = w * (1.0 / torch.clamp(nrm, min=1.0))
        z = h @ w + tp[bn]
        if bns is None:
            return z
        if cfg.batch_norm:
            if training:
                mean = z.mean(0); var = z.var(0, unbiased=False)
            else:
                mean, var = tp[bns + "/moving_mean"], tp[bns + "/moving_variance"]
            z = (z - mean) / torch.sqrt(var + M.BN_EPS) * tp[bns + "/gamma"] + tp[bns + "/beta"]
        r = torch.relu(z)
        if masks is not None:
            r = r * torch.from_numpy(masks[li].astype(np.float64)) / keep
        if li == 0:
            h = r; res = h
        elif li % 2 == 1:
            h = r
        else:
            h = res + r if cfg.residual else r
            res = h


CFGS = [M.Config(64, 2, True, True, True), M.Config(64, 2, True, True, False),
        M.Config(48, 1, False, True, True), M.Config(64, 3, True, False, False),
        M.Config(32, 2, False, False, True)]


@pytest.mark.parametrize("cfg", CFGS)
def test_forward_and_grads_match_autograd(cfg):
    rng = np.random.RandomState(7)
    B = 24
    p = M.init_params(cfg.linear_size, cfg.num_layers, seed=3, batch_norm=cfg.batch_norm, bn="trained")
    # make some weights have norm < 1 so that both clip branches are exercised
    p["linear_model/w4"] = p["linear_model/w4"] * 0.01
    x = rng.standard_normal((B, 32)); t = rng.standard_normal((B, 48))
    keep = 0.5
    masks = [(rng.uniform(size=(B, cfg.linear_size)) < keep).astype(np.uint8) for _ in range(2 * cfg.num_layers + 1)]
    tp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=not k.endswith(("moving_mean", "moving_variance")))
          for k, v in p.items()}
    for training in (False, True):
        y = M.forward(p, x, cfg, training=training, keep_prob=keep, masks=masks if training else None)
        yt = torch_forward(tp, torch.from_numpy(x), cfg, training, masks if training else None, keep)
        np.testing.assert_allclose(y, yt.detach().numpy(), rtol=1e-10, atol=1e-12)
    y, cache = M.forward(p, x, cfg, training=True, keep_prob=keep, masks=masks, want_cache=True)
    grads = M.backward(p, x, t, cfg, cache, y)
    yt = torch_forward(tp, torch.from_numpy(x), cfg, True, masks, keep)
    loss = ((yt - torch.from_numpy(t)) ** 2).mean()
    assert abs(loss.item() - M.loss_fn(y, t)) < 1e-12
    loss.backward()
    for n in M.trainable_names(cfg.linear_size, cfg.num_layers, cfg.batch_norm):
        g_t = tp[n].grad.numpy()
        scale = max(1e-12, np.abs(g_t).max())
        assert np.abs(grads[n] - g_t).max() <= 1e-9 * max(scale, 1e-3), n
    if cfg.batch_norm:      # bias before BN has exactly zero gradient
        assert np.abs(grads["linear_model/b1"]).max() < 1e-12


def test_fold_matches_inference():
    for cfg in CFGS:
        p = M.init_params(cfg.linear_size, cfg.num_layers, seed=5, batch_norm=cfg.batch_norm, bn="trained")
        x = np.random.RandomState(1).standard_normal((10, 32))
        y = M.forward(p, x, cfg, training=False)
        yf = M.forward_folded(M.fold_inference(p, cfg), x, cfg)
        np.testing.assert_allclose(y, yf, rtol=1e-10, atol=1e-12)


def test_adam_is_tf_flavoured_and_matches_torch_rewrite():
    """Three steps of train_step vs. a torch re-statement using autograd gradients and the TF update rule."""
    cfg = M.Config(32, 1, True, True, True)
    p = M.init_params(32, 1, seed=2)
    tp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=not k.endswith(("moving_mean", "moving_variance")))
          for k, v in p.items()}
    st = M.AdamState()
    m = {k: torch.zeros_like(v) for k, v in tp.items()}
    v_ = {k: torch.zeros_like(v) for k, v in tp.items()}
    rng = np.random.RandomState(0)
    lr0 = 1e-3
    for step in range(3):
        x = rng.standard_normal((16, 32)); t = rng.standard_normal((16, 48))
        loss, lr_t, y = M.train_step(p, st, x, t, cfg, lr0)
        yt = torch_forward(tp, torch.from_numpy(x), cfg, True)
        lt = ((yt - torch.from_numpy(t)) ** 2).mean()
        assert abs(lt.item() - loss) < 1e-10
        for k in tp:
            if tp[k].grad is not None:
                tp[k].grad = None
        lt.backward()
        tt = step + 1
        lr = lr0 * 0.96 ** (step / 100000.0)
        assert abs(lr - lr_t) < 1e-15
        alpha = lr * np.sqrt(1 - 0.999 ** tt) / (1 - 0.9 ** tt)
        with torch.no_grad():
            for k in tp:
                if tp[k].grad is None:
                    continue
                g = tp[k].grad
                m[k] += (g - m[k]) * 0.1
                v_[k] += (g * g - v_[k]) * 0.001
                tp[k] -= alpha * m[k] / (torch.sqrt(v_[k]) + 1e-8)
        for k in ("linear_model/w1", "linear_model/w4", "linear_model/b4",
                  "linear_model/batch_normalization/gamma", "linear_model/two_linear_0/w3_0"):
            np.testing.assert_allclose(p[k], tp[k].detach().numpy(), rtol=1e-7, atol=1e-9)
    assert st.t == 3
    # BN moving stats moved towards batch stats with momentum 0.99 and the BIASED variance
    assert not np.allclose(p["linear_model/batch_normalization/moving_mean"], 0)


def test_syncbn_sharding_equivalence():
    """batch statistics summed over row shards == whole-batch statistics."""
    rng = np.random.RandomState(3)
    z = rng.standard_normal((64, 16))
    parts = np.split(z, 4)
    s1 = sum(q.sum(0) for q in parts); s2 = sum((q * q).sum(0) for q in parts)
    mean = s1 / 64; var = s2 / 64 - mean ** 2
    np.testing.assert_allclose(mean, z.mean(0), atol=1e-14)
    np.testing.assert_allclose(var, z.var(0), atol=1e-13)


def test_philox_known_answer_and_mask_rate():
    # Random123 known-answer test vectors for philox4x32-10
    z = [np.zeros((1, 1), np.uint64)] * 4
    out = M.philox4x32_10(z, [np.uint64(0), np.uint64(0)])
    assert [int(o[0, 0]) for o in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = [np.full((1, 1), 0xFFFFFFFF, np.uint64)] * 4
    out = M.philox4x32_10(f, [np.uint64(0xFFFFFFFF), np.uint64(0xFFFFFFFF)])
    assert [int(o[0, 0]) for o in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    m = M.dropout_mask_philox(1234, 7, 2, 256, 128, 0.5)
    assert 0.47 < m.mean() < 0.53
    # shard independence: rows 100..199 of the global mask == mask generated with row0=100
    m2 = M.dropout_mask_philox(1234, 7, 2, 100, 128, 0.5, row0=100)
    assert np.array_equal(m[100:200], m2)
    assert M.dropout_mask_philox(1, 0, 0, 8, 8, 1.0).all()


def test_quant_hook_identity_and_bf16():
    """forward/backward(quant=...) restate the tensor-core path's rounding points: with an identity rounding
    they reproduce the exact graph; with bfloat16 rounding the outputs move by ~2^-9 and the gradients stay
    close in relative L2 (a few ReLU derivatives flip, so the max-norm is NOT a meaningful comparison)."""
    from helpers import bf16_round
    cfg = M.Config(128, 1, True, True, True)
    p = {k: v.astype(np.float64) for k, v in M.init_params(128, 1, seed=3, bn="trained").items()}
    rng = np.random.RandomState(0)
    x, t = rng.standard_normal((64, 32)), rng.standard_normal((64, 48))
    masks = [(rng.uniform(size=(64, 128)) < 0.5).astype(np.uint8) for _ in range(3)]
    y0, c0 = M.forward(p, x, cfg, training=True, keep_prob=0.5, masks=masks, want_cache=True)
    g0 = M.backward(p, x, t, cfg, c0, y0)
    y1, c1 = M.forward(p, x, cfg, training=True, keep_prob=0.5, masks=masks, want_cache=True, quant=lambda a: a)
    g1 = M.backward(p, x, t, cfg, c1, y1, quant=lambda a: a)
    assert np.abs(y1 - y0).max() < 1e-12
    for k in g0:
        if np.abs(g0[k]).max() > 1e-12:          # biases in front of a BN layer: zero up to rounding
            assert np.abs(g1[k] - g0[k]).max() <= 1e-10 * np.abs(g0[k]).max(), k
    q = lambda a: bf16_round(a).astype(np.float64)  # noqa: E731
    y2, c2 = M.forward(p, x, cfg, training=True, keep_prob=0.5, masks=masks, want_cache=True, quant=q)
    g2 = M.backward(p, x, t, cfg, c2, y2, quant=q)
    assert 1e-5 < np.abs(y2 - y0).max() < 2e-2 * np.abs(y0).max()
    for k in g0:
        if np.abs(g0[k]).max() > 1e-12:
            assert np.linalg.norm(g2[k] - g0[k]) < 0.2 * np.linalg.norm(g0[k]), k
