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


def test_summaries_dir_gets_tensorboard_event_files(tmp_path):
    """The logging calls of the reference's train() (predict_3dpose.py:248-253, :322-323) against a model built with a
    summaries_dir (linear_model.py:80-82): the event files hold the step() summaries, step by step."""
    from p3d import LinearModel, summary
    m = LinearModel(256, 2, True, True, True, 64, 1e-3, str(tmp_path / "log"), mode="fp32", seed=3)
    x, t = synth.mlp_inputs(64, seed=5)
    losses, lrs = [], []
    for _ in range(3):
        step_loss, loss_summary, lr_summary, _ = m.step(None, x, t, 0.5, isTraining=True)
        current_step = m.global_step.eval()
        m.train_writer.add_summary(loss_summary, current_step)
        m.train_writer.add_summary(lr_summary, current_step)
        losses.append(np.float32(step_loss)); lrs.append(np.float32(m.learning_rate.eval()))
    m.test_writer.add_summary(m.err_mm_summary(47.25), current_step)
    train_path, test_path = m.train_writer.path, m.test_writer.path
    m.close()
    assert os.path.dirname(train_path) == str(tmp_path / "log" / "train")
    ev = summary.read_events(train_path)[1:]
    assert [e["step"] for e in ev] == [1, 1, 2, 2, 3, 3]
    assert [e["scalars"]["loss/loss"] for e in ev[0::2]] == losses
    got_lr = np.array([e["scalars"]["learning_rate/learning_rate"] for e in ev[1::2]], dtype=np.float32)
    assert np.allclose(got_lr, 1e-3, rtol=1e-4)
    evt = summary.read_events(test_path)
    assert evt[1]["step"] == 3 and evt[1]["scalars"] == {"loss/error_mm": 47.25}
