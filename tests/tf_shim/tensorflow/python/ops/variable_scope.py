"""`from tensorflow.python.ops import variable_scope as vs` (src/linear_model.py:8): vs.variable_scope(...)."""
from tensorflow import variable_scope, get_variable  # noqa: F401
