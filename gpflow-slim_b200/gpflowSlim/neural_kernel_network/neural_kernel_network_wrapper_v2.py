"""v2 wrapper (reference neural_kernel_network_wrapper_v2.py): the same Linear / Product maths
expressed on Python lists of [N, M] matrices in the reference (:116-124, :151-155).  On the
fused path the list formulation is irrelevant -- both versions compile to the same program --
so v2 reuses the v1 layer classes."""
from .neural_kernel_network_wrapper import Activation, Linear, NKNWrapper, Product  # noqa: F401


class NKNWrapperV2(NKNWrapper):
    pass
