"""tensorflow.contrib.eager of the TF-1.x shim (test infrastructure only; see ../__init__.py)."""
import torch


def in_eager_mode():
    return True


def implicit_value_and_gradients(f):
    import tensorflow as tf

    def run():
        val = f()
        vs = [v for v in tf.shim_variables() if v.requires_grad]
        gs = torch.autograd.grad(val, vs, allow_unused=True)
        return val, list(zip(gs, vs))
    return run
