"""CPU oracle for the LinearModel lifting network.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product path (3d-pose-baseline_b200/p3d) never
does and fails loudly when the CUDA library is missing.

PARITY UNPINNED at the TensorFlow boundary: the reference's model lives in
TensorFlow (un-vendored, version unpinned, README.md:23) which is not
installable here, and the reference ships no tests, seeds or golden vectors for
it.  This file restates, in NumPy, the graph that
/root/reference/src/linear_model.py builds, with the TF op semantics spelled
out; it is cross-checked against torch autograd (tests/test_oracle_mlp.py) and
against the reference's TF2 twin `PoseBase.call`
(src/top_vae_3d_pose/models.py:442-481) by reading, not by execution.

Everything is dtype-parametrised: float64 is the parity oracle, float32 is what
bench.py times as the CPU baseline ("port").
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

BN_EPS = 1e-3        # tf.layers.batch_normalization default epsilon
BN_MOMENTUM = 0.99   # tf.layers.batch_normalization default momentum
ADAM_B1, ADAM_B2, ADAM_EPS = 0.9, 0.999, 1e-8   # tf.train.AdamOptimizer defaults
LR_DECAY_STEPS, LR_DECAY_RATE = 100000, 0.96    # linear_model.py:88-90


# --------------------------------------------------------------------------- names
def layer_names(num_layers: int):
    """TF variable names, in graph order, of every (weight, bias, bn-scope) triple.

    linear_model.py:106-107 (w1,b1, "batch_normalization"), :176-193 (w2_i,b2_i,
    "batch_normalization1<i>", w3_i,b3_i,"batch_normalization2<i>"), :121-122 (w4,b4).
    """
    out = [("linear_model/w1", "linear_model/b1", "linear_model/batch_normalization")]
    for i in range(num_layers):
        s = f"linear_model/two_linear_{i}/"
        out.append((s + f"w2_{i}", s + f"b2_{i}", s + f"batch_normalization1{i}"))
        out.append((s + f"w3_{i}", s + f"b3_{i}", s + f"batch_normalization2{i}"))
    out.append(("linear_model/w4", "linear_model/b4", None))
    return out


def param_shapes(linear_size: int, num_layers: int, out_size: int = 48, in_size: int = 32,
                 batch_norm: bool = True):
    """name -> shape for every variable the reference creates (weights are [in,out])."""
    shapes = {}
    names = layer_names(num_layers)
    for li, (w, b, bn) in enumerate(names):
        k = in_size if li == 0 else linear_size
        n = out_size if li == len(names) - 1 else linear_size
        shapes[w] = (k, n)
        shapes[b] = (n,)
        if bn is not None and batch_norm:
            for leaf in ("gamma", "beta", "moving_mean", "moving_variance"):
                shapes[f"{bn}/{leaf}"] = (n,)
    return shapes


def trainable_names(linear_size, num_layers, batch_norm=True):
    return [n for n in param_shapes(linear_size, num_layers, batch_norm=batch_norm)
            if not n.endswith(("moving_mean", "moving_variance"))]


# --------------------------------------------------------------------------- init
def kaiming(shape, rng: np.random.RandomState, dtype=np.float64):
    """linear_model.py:17-29: truncated_normal(shape) * sqrt(2/shape[0]).

    tf.truncated_normal re-draws samples beyond two standard deviations; the
    same distribution is drawn here by rejection.  (For a bias, shape[0] is its
    own length — the reference uses the same initializer for biases.)
    """
    n = int(np.prod(shape))
    vals = rng.standard_normal(n)
    bad = np.abs(vals) > 2.0
    while bad.any():
        vals[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(vals) > 2.0
    return (vals.reshape(shape) * math.sqrt(2.0 / float(shape[0]))).astype(dtype)


def init_params(linear_size=1024, num_layers=2, out_size=48, seed=1, batch_norm=True,
                bn="fresh", dtype=np.float64):
    """Random-init variables.  bn='fresh' = TF defaults (gamma 1, beta 0, moving mean 0,
    moving variance 1); bn='trained' = non-trivial statistics so that folding is exercised
    (SURVEY.md §8d)."""
    rng = np.random.RandomState(seed)
    p = {}
    for name, shape in param_shapes(linear_size, num_layers, out_size, batch_norm=batch_norm).items():
        leaf = name.rsplit("/", 1)[-1]
        if leaf in ("gamma", "beta", "moving_mean", "moving_variance"):
            continue
        p[name] = kaiming(shape, rng, dtype)
    rng2 = np.random.RandomState(seed + 1)
    for name, shape in param_shapes(linear_size, num_layers, out_size, batch_norm=batch_norm).items():
        leaf = name.rsplit("/", 1)[-1]
        if leaf == "gamma":
            p[name] = (np.ones(shape) if bn == "fresh" else rng2.uniform(0.5, 1.5, shape)).astype(dtype)
        elif leaf == "beta":
            p[name] = (np.zeros(shape) if bn == "fresh" else rng2.normal(0, 0.1, shape)).astype(dtype)
        elif leaf == "moving_mean":
            p[name] = (np.zeros(shape) if bn == "fresh" else rng2.normal(0, 0.5, shape)).astype(dtype)
        elif leaf == "moving_variance":
            p[name] = (np.ones(shape) if bn == "fresh" else rng2.uniform(0.5, 2.0, shape)).astype(dtype)
    return p


# --------------------------------------------------------------------------- ops
def clip_by_norm(w, clip=1.0):
    """tf.clip_by_norm(w, 1) with no `axes` (linear_model.py:108,123,178,189):
    whole-tensor L2 norm;  w * clip / max(||w||_F, clip)."""
    nrm = np.sqrt(np.sum(w * w))
    return w * (clip / max(nrm, clip)), nrm


@dataclass
class Config:
    linear_size: int = 1024
    num_layers: int = 2
    residual: bool = True
    batch_norm: bool = True
    max_norm: bool = True
    out_size: int = 48


def forward(p, x, cfg: Config, training=False, keep_prob=1.0, masks=None, want_cache=False, quant=None):
    """The graph of linear_model.py:102-125 + two_linear :154-201.

    quant: None for the reference arithmetic.  A rounding function (e.g. bfloat16 round-to-nearest)
    makes this a statement of the tensor-core training path's rounding points instead: both operands
    of every MatMul are rounded (activations and the UNCLIPPED weights), the clip scale multiplies the
    product, everything else stays in high precision.  Used to test that path at tight tolerances:
    against the exact graph a reduced-precision forward flips the ReLU derivative of the few units
    whose pre-activation is within rounding of zero, which is a large error in max-norm.

    masks: list of 0/1 arrays [B,L] (one per hidden layer, graph order) standing in
    for floor(keep_prob + U[0,1)) of tf.nn.dropout (linear_model.py:114,184,196);
    kept units are scaled by 1/keep_prob.  None = no dropout (keep_prob 1.0).
    Returns y (and the cache the backward pass needs).
    """
    names = layer_names(cfg.num_layers)
    dt = x.dtype
    cache = []
    h = x
    res = None
    n_hidden = len(names) - 1
    for li, (wn, bn_, bns) in enumerate(names):
        w = p[wn]
        nrm = None
        if cfg.max_norm:
            wc, nrm = clip_by_norm(w)
        else:
            wc = w
        if quant is None:
            z = h @ wc + p[bn_]
            h_op, w_op, s_op = h, wc, 1.0
        else:
            h_op, w_op = quant(h).astype(dt), quant(w).astype(dt)
            s_op = (1.0 / max(nrm, 1.0)) if cfg.max_norm else 1.0
            z = (h_op @ w_op) * dt.type(s_op) + p[bn_]
        if li == n_hidden:             # output layer: no BN / ReLU / dropout (linear_model.py:124)
            if want_cache:
                cache.append(dict(h_in=h_op, wc=w_op, nrm=nrm, s=s_op))
            h = z
            break
        c = dict(h_in=h_op, wc=w_op, nrm=nrm, s=s_op)
        if cfg.batch_norm:
            g, b = p[bns + "/gamma"], p[bns + "/beta"]
            if training:
                mean = z.mean(axis=0)
                var = ((z - mean) ** 2).mean(axis=0)          # biased, as TF's non-fused path
            else:
                mean, var = p[bns + "/moving_mean"], p[bns + "/moving_variance"]
            rstd = 1.0 / np.sqrt(var + dt.type(BN_EPS))
            xhat = (z - mean) * rstd
            a = xhat * g + b
            c.update(xhat=xhat, rstd=rstd, mean=mean, var=var)
        else:
            a = z
        r = np.maximum(a, 0)
        c["relu_mask"] = a > 0
        if masks is not None:
            m = masks[li].astype(dt) / dt.type(keep_prob)
            r = r * m
            c["drop"] = m
        # residual: input layer feeds block 0; each block = two hidden layers
        if li == 0:
            h = r
            res = h
        elif li % 2 == 1:              # first linear of a block
            h = r
        else:                          # second linear of a block (linear_model.py:199)
            h = (res + r) if cfg.residual else r
            res = h
        cache.append(c)
    return (h, cache) if want_cache else h


def loss_fn(y, t):
    """linear_model.py:129: mean over ALL B*out elements."""
    d = y - t
    return float(np.mean(d * d))


def backward(p, x, t, cfg: Config, cache, y, quant=None):
    """Gradients of loss wrt every trainable variable (what opt.compute_gradients builds,
    linear_model.py:143).  Differentiates through clip_by_norm and batch-stat BN.
    quant: see forward() - the upstream gradient operand of every MatMul is rounded as well."""
    names = layer_names(cfg.num_layers)
    B = x.shape[0]
    grads = {}
    dy = 2.0 * (y - t) / (B * y.shape[1])
    n_hidden = len(names) - 1

    def wgrad(li, dz):
        wn, bn_, _ = names[li]
        c = cache[li]
        dz_op = dz if quant is None else quant(dz).astype(dz.dtype)
        g_wc = c["h_in"].T @ dz_op
        if cfg.max_norm and c["nrm"] is not None and c["nrm"] > 1.0:
            w = p[wn]
            nrm = c["nrm"]
            what = w / nrm
            g_w = (g_wc - what * np.sum(what * g_wc)) / nrm
        else:
            g_w = g_wc
        grads[wn] = g_w
        grads[bn_] = dz.sum(axis=0)
        return (dz_op @ c["wc"].T) * c["s"]

    dh = wgrad(n_hidden, dy)
    dres = None  # gradient flowing along the residual stream into the previous block output
    for li in range(n_hidden - 1, -1, -1):
        c = cache[li]
        if li >= 1 and li % 2 == 0 and cfg.residual:
            # h = res + r : dh is d(h); splits to the skip path and to r
            dres = dh
        dr = dh
        if "drop" in c:
            dr = dr * c["drop"]
        da = dr * c["relu_mask"]
        if cfg.batch_norm:
            _, _, bns = names[li]
            g = p[bns + "/gamma"]
            grads[bns + "/gamma"] = np.sum(da * c["xhat"], axis=0)
            grads[bns + "/beta"] = np.sum(da, axis=0)
            dxhat = da * g
            # batch-statistics BN backward
            dz = c["rstd"] * (dxhat - dxhat.mean(axis=0) - c["xhat"] * (dxhat * c["xhat"]).mean(axis=0))
        else:
            dz = da
        dh = wgrad(li, dz)
        if li >= 1 and li % 2 == 1 and cfg.residual:
            dh = dh + dres          # block input receives skip-path gradient
            dres = None
    return grads


@dataclass
class AdamState:
    m: dict = field(default_factory=dict)
    v: dict = field(default_factory=dict)
    t: int = 0                 # number of apply_gradients calls so far == global_step


def learning_rate_at(lr0, global_step, dtype=np.float64):
    """tf.train.exponential_decay, non-staircase (linear_model.py:86-90)."""
    return dtype(lr0) * dtype(LR_DECAY_RATE) ** (dtype(global_step) / dtype(LR_DECAY_STEPS))


def adam_update(p, grads, st: AdamState, lr0):
    """tf.train.AdamOptimizer(lr) TF formulation (linear_model.py:137-145):
    alpha_t = lr_t*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1); v += (g^2-v)(1-b2);
    theta -= alpha_t * m / (sqrt(v)+eps).  lr_t uses global_step BEFORE the increment."""
    lr_t = learning_rate_at(lr0, st.t)
    t = st.t + 1
    alpha = lr_t * math.sqrt(1.0 - ADAM_B2 ** t) / (1.0 - ADAM_B1 ** t)
    for n, g in grads.items():
        if n not in st.m:
            st.m[n] = np.zeros_like(p[n])
            st.v[n] = np.zeros_like(p[n])
        st.m[n] = st.m[n] + (g - st.m[n]) * (1.0 - ADAM_B1)
        st.v[n] = st.v[n] + (g * g - st.v[n]) * (1.0 - ADAM_B2)
        p[n] = p[n] - alpha * st.m[n] / (np.sqrt(st.v[n]) + ADAM_EPS)
    st.t = t
    return lr_t


def train_step(p, st: AdamState, x, t, cfg: Config, lr0, keep_prob=1.0, masks=None):
    """One `model.step(..., isTraining=True)` (linear_model.py:203-235): forward with batch
    statistics, loss, gradients, BN moving-average updates (UPDATE_OPS tied to the train op,
    linear_model.py:138-145), Adam.  Returns (loss, lr used, outputs) — outputs are the
    forward outputs computed BEFORE the update, as session.run returns them."""
    y, cache = forward(p, x, cfg, training=True, keep_prob=keep_prob, masks=masks, want_cache=True)
    loss = loss_fn(y, t)
    grads = backward(p, x, t, cfg, cache, y)
    if cfg.batch_norm:
        names = layer_names(cfg.num_layers)
        for li, (_, _, bns) in enumerate(names[:-1]):
            c = cache[li]
            mm, mv = bns + "/moving_mean", bns + "/moving_variance"
            p[mm] = p[mm] * BN_MOMENTUM + c["mean"] * (1.0 - BN_MOMENTUM)
            p[mv] = p[mv] * BN_MOMENTUM + c["var"] * (1.0 - BN_MOMENTUM)
    lr_t = adam_update(p, grads, st, lr0)
    return loss, lr_t, y


# --------------------------------------------------------------------------- inference folding
def fold_inference(p, cfg: Config, dtype=np.float64):
    """What the CUDA weight-prep kernel must produce: per layer W' [K,N] and b' [N] with
    clip_by_norm and the moving-statistics BN affine folded in:
      s = gamma/sqrt(mv+eps);  W' = clip(W)*s;  b' = (b-mm)*s+beta."""
    out = []
    names = layer_names(cfg.num_layers)
    for li, (wn, bn_, bns) in enumerate(names):
        w = p[wn].astype(dtype)
        wc = clip_by_norm(w)[0] if cfg.max_norm else w
        b = p[bn_].astype(dtype)
        if bns is not None and cfg.batch_norm:
            s = p[bns + "/gamma"] / np.sqrt(p[bns + "/moving_variance"] + BN_EPS)
            wc = wc * s
            b = (b - p[bns + "/moving_mean"]) * s + p[bns + "/beta"]
        out.append((wc, b))
    return out


def forward_folded(folded, x, cfg: Config):
    """Inference through folded weights (same value as forward(training=False))."""
    h = x
    res = None
    last = len(folded) - 1
    for li, (w, b) in enumerate(folded):
        z = h @ w + b
        if li == last:
            return z
        r = np.maximum(z, 0)
        if li == 0:
            h = r; res = h
        elif li % 2 == 1:
            h = r
        else:
            h = (res + r) if cfg.residual else r
            res = h
    return h


def dropout_mask_philox(seed, step, layer, rows, cols, keep_prob, row0=0):
    """The documented counter-based dropout mask shared by oracle and CUDA kernel:
    Philox4x32-10 with key=(seed_lo, seed_hi), counter=(global_row, col//4, layer, step);
    word (col%4) -> u = word * 2^-32;  keep iff floor(keep_prob + u) == 1, i.e. u >= 1-keep_prob.
    Rows are GLOBAL rows so the mask does not depend on how the batch is sharded."""
    r = np.arange(row0, row0 + rows, dtype=np.uint64)[:, None]
    c4 = (np.arange(cols, dtype=np.uint64) // 4)[None, :]
    ctr = [np.broadcast_to(r, (rows, cols)).astype(np.uint64),
           np.broadcast_to(c4, (rows, cols)).astype(np.uint64),
           np.full((rows, cols), layer, np.uint64),
           np.full((rows, cols), step & 0xFFFFFFFF, np.uint64)]
    key = [np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)]
    words = philox4x32_10(ctr, key)
    sel = (np.arange(cols) % 4)[None, :]
    w = np.choose(np.broadcast_to(sel, (rows, cols)), words)
    u = w.astype(np.float64) * (1.0 / 4294967296.0)
    u32 = u.astype(np.float32)           # the kernel forms u in fp32: word * 2^-32 rounded
    return (np.floor(np.float32(keep_prob) + u32) >= 1.0).astype(np.uint8)


def philox4x32_10(ctr, key):
    """Philox4x32-10 (Salmon et al., SC'11) on uint64-held 32-bit lanes."""
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    W0, W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
    mask = np.uint64(0xFFFFFFFF)
    c0, c1, c2, c3 = [c.copy() for c in ctr]
    k0, k1 = key
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        k0 = (k0 + W0) & mask
        k1 = (k1 + W1) & mask
    return [c0, c1, c2, c3]
