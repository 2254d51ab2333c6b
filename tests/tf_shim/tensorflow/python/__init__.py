"""tensorflow.python (shim): only the sub-package path the reference imports from."""
