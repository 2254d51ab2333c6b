"""Known-answer checks of the TensorFlow op stand-in (tests/tf_shim) that oracle/make_golden_mlp.py runs the
reference's model code on: each op against hand-computed values of the formula TensorFlow documents for it."""
import os
import sys

import numpy as np

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tf_shim")
sys.path.insert(0, _SHIM)
try:
    import tensorflow as tf  # noqa: E402  (the shim)
finally:
    # keep the stand-in out of everybody else's way: other packages (tensorboard) probe `import tensorflow`
    sys.path.remove(_SHIM)
    for _m in [m for m in sys.modules if m == "tensorflow" or m.startswith("tensorflow.")]:
        del sys.modules[_m]


def setup_function(_):
    tf.reset_default_graph()


def test_clip_by_norm_whole_tensor_and_gradient():
    w = tf.Variable(np.array([[3.0, 0.0], [0.0, 4.0]]), dtype=tf.float64, name="w")      # ||w||_F = 5
    c = tf.clip_by_norm(w, 1)
    np.testing.assert_allclose(c.eval(), [[0.6, 0.0], [0.0, 0.8]], atol=1e-15)
    small = tf.Variable(np.array([[0.3, 0.0], [0.0, 0.4]]), dtype=tf.float64, name="s")   # norm 0.5 < 1: unchanged
    np.testing.assert_allclose(tf.clip_by_norm(small, 1).eval(), small.eval(), atol=0)
    # d/dw sum(clip(w) * G) = (G - what <what, G>) / ||w||
    G = np.array([[1.0, 2.0], [3.0, 4.0]])
    loss = tf.reduce_mean(c * G) * 4.0
    opt = tf.train.AdamOptimizer(0.1)
    (g, _), = opt.compute_gradients(loss, [w])
    what = w.eval() / 5.0
    np.testing.assert_allclose(g.eval(), (G - what * (what * G).sum()) / 5.0, atol=1e-15)


def test_batch_normalization_training_and_inference():
    x = tf.placeholder(tf.float64, [None, 2])
    training = tf.placeholder(tf.bool)
    y = tf.layers.batch_normalization(x, training=training, name="bn")
    v = {u.name: u for u in tf.global_variables()}
    assert sorted(v) == ["bn/beta", "bn/gamma", "bn/moving_mean", "bn/moving_variance"]
    assert [u.name for u in tf.trainable_variables()] == ["bn/gamma", "bn/beta"]
    data = np.array([[1.0, 10.0], [3.0, 30.0]])
    sess = tf.Session()
    out = sess.run(y, {x: data, training: True})
    mean, var = data.mean(0), data.var(0)                                  # biased variance
    np.testing.assert_allclose(out, (data - mean) / np.sqrt(var + 1e-3), atol=1e-12)
    np.testing.assert_allclose(v["bn/moving_mean"].eval(), 0.0)            # nothing depends on the update ops yet
    ups = tf.get_collection(tf.GraphKeys.UPDATE_OPS)
    assert len(ups) == 1 or len(ups) == 2
    sess.run(ups, {x: data, training: True})
    np.testing.assert_allclose(v["bn/moving_mean"].eval(), 0.01 * mean, atol=1e-15)
    np.testing.assert_allclose(v["bn/moving_variance"].eval(), 0.99 + 0.01 * var, atol=1e-15)
    out = sess.run(y, {x: data, training: False})
    np.testing.assert_allclose(out, (data - 0.01 * mean) / np.sqrt(0.99 + 0.01 * var + 1e-3), atol=1e-12)


def test_dropout_is_the_keep_prob_form():
    x = tf.placeholder(tf.float64, [None, 4])
    kp = tf.placeholder(tf.float32)
    y = tf.nn.dropout(x, kp)
    data = np.arange(8.0).reshape(2, 4) + 1
    tf.set_dropout_masks([np.array([[1, 0, 1, 0], [0, 0, 1, 1]])])
    np.testing.assert_allclose(tf.Session().run(y, {x: data, kp: 0.5}), data * 2 * [[1, 0, 1, 0], [0, 0, 1, 1]])
    tf.set_dropout_masks(None)
    big = np.ones((200, 4))
    out = tf.Session().run(y, {x: big, kp: 0.25})
    assert set(np.unique(out)) <= {0.0, 4.0} and 0.15 < (out > 0).mean() < 0.35
    np.testing.assert_allclose(tf.Session().run(y, {x: big, kp: 1.0}), big)      # floor(1 + U) == 1


def test_adam_first_steps_and_learning_rate_decay():
    w = tf.Variable(np.array([1.0, -2.0]), dtype=tf.float64, name="w")
    step = tf.Variable(0, trainable=False, name="global_step")
    lr = tf.train.exponential_decay(tf.Variable(0.1, trainable=False, dtype=tf.float64, name="lr"), step, 100000, 0.96)
    loss = tf.reduce_mean(tf.square(w))                                          # d/dw = w
    opt = tf.train.AdamOptimizer(lr)
    train = opt.apply_gradients(opt.compute_gradients(loss), global_step=step)
    sess = tf.Session()
    ref, m, v = np.array([1.0, -2.0]), np.zeros(2), np.zeros(2)
    for t in range(1, 4):
        g = ref.copy()
        lr_t = 0.1 * 0.96 ** ((t - 1) / 100000.0) * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        m += (g - m) * 0.1
        v += (g * g - v) * 0.001
        ref -= lr_t * m / (np.sqrt(v) + 1e-8)
        sess.run(train)
        np.testing.assert_allclose(w.eval(), ref, rtol=1e-13)
    assert int(step.eval()) == 3
    np.testing.assert_allclose(lr.eval(), 0.1 * 0.96 ** (3 / 100000.0), rtol=1e-15)


def test_one_run_shares_one_forward_pass():
    """loss fetched next to the train op is the PRE-update loss (one graph execution), as in session.run."""
    w = tf.Variable(np.array([2.0]), dtype=tf.float64, name="w")
    loss = tf.reduce_mean(tf.square(w))
    opt = tf.train.AdamOptimizer(0.5)
    train = opt.apply_gradients(opt.compute_gradients(loss))
    _, l0 = tf.Session().run([train, loss])
    assert l0 == 4.0 and w.eval()[0] < 2.0


def test_truncated_normal_stays_within_two_sigma():
    tf.set_random_seed(3)
    v = tf.truncated_normal([20000], dtype=tf.float64).eval()
    assert np.abs(v).max() <= 2.0 and 0.86 < v.std() < 0.90          # std of N(0,1) truncated at 2 sigma = 0.8796
