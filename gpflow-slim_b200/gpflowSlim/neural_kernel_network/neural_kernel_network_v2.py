"""NeuralKernelNetworkV2 (reference neural_kernel_network_v2.py:25-40)."""
from .neural_kernel_network import NeuralKernelNetwork


class NeuralKernelNetworkV2(NeuralKernelNetwork):
    pass
