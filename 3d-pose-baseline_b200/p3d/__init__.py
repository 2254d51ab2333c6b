"""p3d - B200-native hot path of 3d-pose-baseline.

Module names mirror the reference's src/ layout so that callers switch by changing the import:
    linear_model.LinearModel, cameras, data_utils, procrustes  (+ evaluate for the MPJPE arithmetic of
    predict_3dpose.evaluate_batches, realtime for the frame loop of openpose_3dpose_sandbox_realtime.py,
    checkpoint for tf.train.Saver's file format, summary for tf.summary.FileWriter's event files).
All compute goes through libp3d.so (sm_100a CUDA); importing this package fails if it is not built.
"""
from . import _lib                      # noqa: F401  (raises ImportError when libp3d.so is missing)
from . import cameras, checkpoint, data_utils, evaluate, linear_model, procrustes, realtime, summary   # noqa: F401
from .linear_model import LinearModel   # noqa: F401

__all__ = ["LinearModel", "cameras", "data_utils", "procrustes", "evaluate", "linear_model", "realtime", "checkpoint", "summary"]
