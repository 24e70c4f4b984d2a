"""NeuralKernelNetwork (reference neural_kernel_network.py:25-47)."""
import torch

from .._backend import lib as _lib
from ..kernels import Kernel


class NeuralKernelNetwork(Kernel):
    def __init__(self, input_dim, primitive_kernels, nknWrapper):
        super().__init__(input_dim)
        self._primitive_kernels = primitive_kernels
        self._nknWrapper = nknWrapper
        self._parameters = self._parameters + self._nknWrapper.parameters
        for kern in self._primitive_kernels:
            self._parameters = self._parameters + kern.parameters

    def _emit(self, b, presliced=False):
        if not self._nknWrapper.fusable:
            raise NotImplementedError('Activation layers are evaluated outside the fused kernel')
        refs = [k._emit(b, presliced) for k in self._primitive_kernels]
        # Linear / Product read a contiguous slot range: plain primitives emitted back to back
        # already are one; anything else is gathered with COPY ops.
        contiguous = all(r[0] == 'p' for r in refs) and \
            all(refs[i + 1][1] == refs[i][1] + 1 for i in range(len(refs) - 1))
        if not contiguous:
            refs = [b.op(_lib.GPS_OP_COPY, r) for r in refs]
        src, width = refs[0], len(refs)
        for layer in self._nknWrapper.layers:
            assert layer.input_dim == width, 'NKN layer width mismatch'
            src, width = layer.emit(b, src)
        assert width == 1, 'output of NKN must only have one term'
        return src

    def _unfused(self, vals, shape):
        h = torch.stack([v.reshape(-1) for v in vals], 1)
        return self._nknWrapper.forward(h).reshape(shape)

    # `fusable` is False for Activation layers, for composed primitive kernels (kernels.py: RatQuad,
    # Polynomial, ...) and for networks too large for the fused kernel's tables: all take the
    # stacked-Gram route (neural_kernel_network.py:35-47 of the reference, on the device)
    def _Kdiag_composed(self, X, presliced):
        vals = [k.Kdiag(X, presliced) for k in self._primitive_kernels]
        return self._unfused(vals, vals[0].shape)

    def _K_composed(self, X, X2, presliced):
        vals = [k.K(X, X2, presliced) for k in self._primitive_kernels]
        return self._unfused(vals, vals[0].shape)
