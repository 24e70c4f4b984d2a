"""tf.contrib namespace of the TF-1.x shim (test infrastructure only; see ../__init__.py)."""
