"""NKN layer stack (the surface of the reference's neural_kernel_network_wrapper.py:29-173:
`NKNWrapper(hparams)`, layers `Linear`, `Product`, `Activation`, hparams = list of
{'name', 'params'}).

On the B200 path a layer is first of all an EMITTER: `NeuralKernelNetwork` compiles Linear /
Product layers into GPS_OP_LINEAR / GPS_OP_PRODUCT ops of the fused Gram program, so the
[N*M, k] activations of the reference (wrapper.py:42-47) never exist in memory.  Everything
else a layer can do -- the dense `forward` used by the unfused fallback (Activation layers,
composed primitive kernels) and the sympy pretty-printer `symbolic` -- is derived from ONE
method, `combine(channels)`, which maps a list of per-channel values to a list of per-channel
values and does not care whether a value is a tensor column or a sympy symbol."""
import numpy as np
import torch

from .._backend import lib as _lib
from ..params import Parameter
from ..transforms import positive


class _Layer(object):
    """input_dim channels in, output_dim channels out."""
    fusable = True
    parameters = ()

    def __init__(self, input_dim, output_dim, name):
        self.input_dim, self.output_dim, self.name = input_dim, output_dim, name

    def combine(self, channels, numeric=True):
        raise NotImplementedError

    def emit(self, b, src):
        """Append this layer to the fused program; `src` = reference of the first input slot.
        Returns (reference of the first output slot, number of outputs)."""
        raise NotImplementedError('%s cannot be fused' % type(self).__name__)

    # -- derived ------------------------------------------------------------------------
    def forward(self, input):                      # [nm, input_dim] -> [nm, output_dim]
        return torch.stack(self.combine(list(input.unbind(1))), 1)

    def symbolic(self, ks):
        return self.combine(list(ks), numeric=False)


class Linear(_Layer):
    """out_i = b_i + sum_j W_ij in_j with W, b > 0 (wrapper.py:90-129).  W is drawn from numpy's
    GLOBAL RNG, U(1/(2 in), 3/(2 in)), b = 0.01 -- like the reference (:100-104), so seeding numpy
    reproduces its initial state."""

    def __init__(self, input_dim, output_dim, name='Linear'):
        super().__init__(input_dim, output_dim, name)
        lo, hi = 0.5 / input_dim, 1.5 / input_dim
        w0 = np.random.uniform(low=lo, high=hi, size=[output_dim, input_dim]).astype(np.float64)
        self._weights = Parameter(w0, transform=positive, name='weights')
        self._bias = Parameter(np.full([output_dim], 0.01), transform=positive, name='bias')
        self.parameters = [self._weights, self._bias]

    weights = property(lambda self: self._weights.value)
    bias = property(lambda self: self._bias.value)

    def forward(self, input):                      # dense fast path of the generic rule
        return input @ self.weights.t() + self.bias

    def combine(self, channels, numeric=True):
        W, b = self.weights, self.bias
        if not numeric:
            W, b = W.detach().cpu().numpy(), b.detach().cpu().numpy()
        out = []
        for i in range(self.output_dim):
            acc = b[i]
            for j in range(self.input_dim):
                acc = acc + channels[j] * W[i, j]
            out.append(acc)
        return out

    def emit(self, b, src):
        w_off = b.theta(lambda: self.weights, self.output_dim * self.input_dim)
        b_off = b.theta(lambda: self.bias, self.output_dim)
        return b.op(_lib.GPS_OP_LINEAR, src, self.input_dim, w_off, b_off, n=self.output_dim), self.output_dim


class Product(_Layer):
    """Products of consecutive groups of `step` channels (wrapper.py:132-155)."""

    def __init__(self, input_dim, step, name='Product'):
        if not (isinstance(step, int) and step > 1):
            raise AssertionError('step must be number greater than 1')
        if input_dim % step:
            raise AssertionError('input dim must be multiples of step')
        super().__init__(input_dim, input_dim // step, name)
        self.step = step

    def combine(self, channels, numeric=True):
        out = []
        for g in range(self.output_dim):
            acc = channels[g * self.step]
            for c in channels[g * self.step + 1:(g + 1) * self.step]:
                acc = acc * c
            out.append(acc)
        return out

    def emit(self, b, src):
        return b.op(_lib.GPS_OP_PRODUCT, src, self.step, n=self.output_dim), self.output_dim


class Activation(_Layer):
    """Arbitrary elementwise callable with its own parameter list (wrapper.py:158-173): cannot be
    fused, evaluated by torch on the stacked primitive Grams."""
    fusable = False

    def __init__(self, input_dim, activation_fn, activation_fn_params, name='Activation'):
        super().__init__(input_dim, input_dim, name)
        self.activation_fn = activation_fn
        self.parameters = activation_fn_params

    def forward(self, input):
        return self.activation_fn(input)

    def combine(self, channels, numeric=True):
        return [self.activation_fn(c) for c in channels]


LAYER_TYPES = dict(Linear=Linear, Product=Product, Activation=Activation)


class NKNWrapper(object):
    def __init__(self, hparams):
        self._layers = [LAYER_TYPES[spec['name']](**spec['params']) for spec in hparams]

    layers = property(lambda self: self._layers)
    fusable = property(lambda self: all(layer.fusable for layer in self._layers))

    @property
    def parameters(self):
        return [p for layer in self._layers for p in layer.parameters]

    def forward(self, input):
        """[nm, k] stacked primitive Grams -> [nm, 1] (the unfused route)."""
        for layer in self._layers:
            input = layer.forward(input)
        return input

    def symbolic(self):
        """sympy expression of the network over symbols k0, k1, ... (wrapper.py:56-61)."""
        import sympy as sp
        ks = sp.symbols(['k' + str(i) for i in range(self._layers[0].input_dim)]) + [1.]
        for layer in self._layers:
            ks = layer.symbolic(ks)
        assert len(ks) == 1, 'output of NKN must only have one term'
        return ks[0]
