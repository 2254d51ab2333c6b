"""The CUDA evaluation path against tests/golden/driver.npz: numbers the reference's own `evaluate_batches`
(src/predict_3dpose.py:352-444, executed line for line by oracle/make_golden_driver.py with recorded model outputs in
place of the TensorFlow session call) returned.  Tolerance (north_star): Procrustes-aligned MPJPE within 1e-3 mm."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def drv(golden_dir):
    return np.load(os.path.join(golden_dir, "driver.npz"))


def _case(drv, p14):
    tag = "p14" if p14 else "p17"
    width = 42 if p14 else 48
    dec = drv["evb_%s_dec" % tag]            # [NB,B,width] float64 (what get_all_batches hands over)
    pred = drv["evb_%s_pred" % tag]          # [NB,B,width] float32 (what session.run returned)
    assert dec.shape[2] == width and pred.dtype == np.float32
    return tag, dec, pred


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("p14", [False, True])
def test_mpjpe_kernel_against_executed_reference(drv, p14, precision):
    from p3d import evaluate
    tag, dec, pred = _case(drv, p14)
    width = dec.shape[2]
    for use_proc in (False, True):
        pt = "proc" if use_proc else "plain"
        tot, joint = evaluate.mpjpe(pred.reshape(-1, width), dec.reshape(-1, width), drv["mean3d"], drv["std3d"],
                                    procrustes=use_proc, predict_14=p14, precision=precision)
        assert abs(tot - float(drv["evb_%s_%s_total" % (tag, pt)])) < 1e-3
        assert np.abs(joint - drv["evb_%s_%s_joint" % (tag, pt)]).max() < 1e-3


class _RecordedModel:
    """What evaluate_batches needs of a model, with the recorded predictions behind `step` (the reference side of the
    golden file had the same stub in place of TensorFlow)."""

    def __init__(self, pred, losses, batch_size, predict_14):
        self.device, self.batch_size, self.predict_14 = torch.cuda.current_device(), batch_size, predict_14
        self._pred, self._losses = pred, losses

    def step(self, sess, enc, dec, keep_prob, isTraining=True):
        assert keep_prob == 1.0 and isTraining is False and enc.is_cuda and dec.is_cuda
        assert enc.shape[0] == self._pred.shape[0] * self._pred.shape[1] == dec.shape[0]
        poses3d = torch.from_numpy(self._pred.reshape(-1, self._pred.shape[2])).to(enc.device)
        return float(np.mean(self._losses)), None, poses3d


@pytest.mark.parametrize("p14", [False, True])
def test_evaluate_batches_against_executed_reference(drv, p14):
    """Same call, same lists of batches, same model outputs as the reference run -> same (total_err, joint_err, loss)."""
    from p3d import evaluate
    tag, dec, pred = _case(drv, p14)
    NB, B = dec.shape[0], dec.shape[1]
    enc = [drv["evb_enc"][i] for i in range(NB)]
    decs = [dec[i] for i in range(NB)]
    model = _RecordedModel(pred, drv["evb_%s_losses" % tag], B, p14)
    for use_proc in (False, True):
        pt = "proc" if use_proc else "plain"
        tot, joint, step_time, loss = evaluate.evaluate_batches(
            None, model, drv["mean3d"], drv["std3d"], drv["evb_%s_use3d" % tag], drv["evb_%s_ignore3d" % tag],
            drv["mean2d"], drv["std2d"], drv["use2d"], drv["ignore2d"], 0, enc, decs, procrustes=use_proc)
        assert abs(tot - float(drv["evb_%s_%s_total" % (tag, pt)])) < 1e-3
        assert joint.shape == (14 if p14 else 17,) and np.abs(joint - drv["evb_%s_%s_joint" % (tag, pt)]).max() < 1e-3
        assert abs(loss - float(drv["evb_%s_%s_loss" % (tag, pt)])) < 1e-6 and step_time > 0
