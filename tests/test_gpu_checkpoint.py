"""model.saver (tf.train.Saver stand-in, src/linear_model.py:151; predict_3dpose.py:165-181,328): a trained model
written as a TensorFlow-format checkpoint and restored into a fresh model continues bit-identically."""
import os

import numpy as np
import pytest

from helpers import make_model
from oracle import mlp_ref as M, synth
from p3d import checkpoint as ck

pytestmark = pytest.mark.gpu


def test_saver_roundtrip_continues_training(tmp_path):
    cfg = M.Config(256, 2, True, True, True)
    a, _ = make_model(cfg, seed=2, mode="fp32", lr=1e-3)
    x, t = synth.mlp_inputs(64, seed=9)
    nh = 5
    masks = [(np.random.RandomState(s).uniform(size=(nh, 64, 256)) < 0.5).astype(np.uint8) for s in range(4)]
    for s in range(2):
        a.step(None, x, t, 0.5, isTraining=True, dropout_mask=masks[s])
    prefix = a.saver.save(None, os.path.join(str(tmp_path), "checkpoint"), global_step=a.global_step.eval())
    assert prefix.endswith("checkpoint-2") and os.path.isfile(prefix + ".index")
    st = ck.get_checkpoint_state(str(tmp_path))
    assert st["model_checkpoint_path"] == prefix
    stored = ck.read_bundle(prefix)
    assert stored["global_step"].dtype == np.int32 and int(stored["global_step"]) == 2
    assert stored["linear_model/two_linear_1/w3_1"].shape == (256, 256)
    assert "linear_model/w4/Adam_1" in stored and "beta1_power" in stored
    assert np.isclose(float(stored["beta2_power"]), 0.999 ** 3)

    b, _ = make_model(cfg, seed=77, mode="fp32", lr=5e-2)            # different init and learning rate
    b.saver.restore(None, st["model_checkpoint_path"])
    assert b.global_step.eval() == 2 and np.isclose(b.learning_rate.eval(), a.learning_rate.eval())
    va, vb = a.get_variables(include_optimizer=True), b.get_variables(include_optimizer=True)
    for k in va:
        assert np.array_equal(va[k], vb[k]), k
    for s in range(2, 4):                                            # both continue identically
        la = a.step(None, x, t, 0.5, isTraining=True, dropout_mask=masks[s])
        lb = b.step(None, x, t, 0.5, isTraining=True, dropout_mask=masks[s])
        assert la[0] == lb[0] and np.array_equal(la[3], lb[3])
    ya = a.step(None, x, t, 1.0, isTraining=False)[2]
    yb = b.step(None, x, t, 1.0, isTraining=False)[2]
    assert np.array_equal(ya, yb)

    with pytest.raises(ValueError):
        b.saver.restore(None, os.path.join(str(tmp_path), "checkpoint-999"))
    c, _ = make_model(M.Config(128, 2, True, True, True), seed=1, mode="fp32")
    with pytest.raises(ValueError, match="shape"):
        c.saver.restore(None, prefix)
    for m in (a, b, c):
        m.close()
