"""tensorflow.python.ops (shim)."""
