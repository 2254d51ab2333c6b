"""tf.keras (shim): the two symbols `PoseBase` (src/top_vae_3d_pose/models.py:287-481) needs - a Model base class whose
instances are called like functions, and layers.BatchNormalization (TF2 defaults: momentum 0.99, epsilon 1e-3; in eager
mode the layer updates its moving statistics inside call() when training=True)."""
import tensorflow as _tf


class Model:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self.call(*a, **k)


class _BatchNormalization:
    def __init__(self, name=None, momentum=0.99, epsilon=1e-3, **kw):
        self._bn = _tf._BatchNorm((_tf._scope_prefix() + (name or "batch_normalization")), momentum, epsilon)

    def __call__(self, x, training=False):
        if not self._bn.built:            # variables are named by the scope the LAYER was created in
            saved, _tf._G.scopes = _tf._G.scopes, []
            try:
                self._bn.build(x._shape[-1], x.dtype or _tf.float32)
            finally:
                _tf._G.scopes = saved
        return self._bn.apply(x, training, graph_mode=False)

    @property
    def variables(self):
        b = self._bn
        return [b.gamma, b.beta, b.moving_mean, b.moving_variance]


class layers:
    BatchNormalization = _BatchNormalization
    Layer = object
