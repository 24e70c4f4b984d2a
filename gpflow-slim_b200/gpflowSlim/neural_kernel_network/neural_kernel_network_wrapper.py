"""NKN layer stack (reference neural_kernel_network_wrapper.py:29-173).

Layers only DESCRIBE the composition: `NeuralKernelNetwork` compiles Linear / Product layers
into GPS_OP_LINEAR / GPS_OP_PRODUCT ops of the fused Gram program, so the [N*M, k]
activations of the reference (wrapper.py:42-47) never exist in memory.  `forward` is kept for
API compatibility (and for `Activation` layers, whose arbitrary callables cannot be fused)."""
import math

import numpy as np
import torch

from .._backend import lib as _lib
from ..params import Parameter
from ..transforms import positive


class NKNWrapper(object):
    def __init__(self, hparams):
        self._LAYERS = dict(Linear=Linear, Product=Product, Activation=Activation)
        self._layers = [self._LAYERS[l['name']](**l['params']) for l in hparams]

    def forward(self, input):
        outputs = input  # [nm, k]
        for l in self._layers:
            outputs = l.forward(outputs)
        return outputs

    @property
    def layers(self):
        return self._layers

    @property
    def fusable(self):
        return all(l.fusable for l in self._layers)

    @property
    def parameters(self):
        params = []
        for l in self._layers:
            params = params + l.parameters
        return params

    def symbolic(self):
        import sympy as sp
        ks = sp.symbols(['k' + str(i) for i in range(self._layers[0].input_dim)]) + [1.]
        for l in self._layers:
            ks = l.symbolic(ks)
        assert len(ks) == 1, 'output of NKN must only have one term'
        return ks[0]


class _KernelLayer(object):
    fusable = True

    def __init__(self, input_dim, name):
        self.input_dim = input_dim
        self.name = name

    def forward(self, input):
        raise NotImplementedError

    @property
    def parameters(self):
        raise NotImplementedError

    def emit(self, b, src):
        """Append this layer to the fused program; `src` = reference of the first input slot."""
        raise NotImplementedError


class Linear(_KernelLayer):
    """y = x W^T + b with positive W, b (wrapper.py:90-129).  W is drawn from numpy's GLOBAL
    RNG, U(1/(2 in), 3/(2 in)), bias 0.01 -- exactly as the reference (wrapper.py:100-104)."""

    def __init__(self, input_dim, output_dim, name='Linear'):
        super().__init__(input_dim, name=name)
        self.output_dim = output_dim
        min_w, max_w = 1. / (2 * input_dim), 3. / (2 * input_dim)
        weights = np.random.uniform(low=min_w, high=max_w, size=[output_dim, input_dim]).astype(np.float64)
        self._weights = Parameter(weights, transform=positive, name='weights')
        self._bias = Parameter(0.01 * np.ones([self.output_dim], dtype=np.float64), transform=positive,
                               name='bias')

    @property
    def weights(self):
        return self._weights.value

    @property
    def bias(self):
        return self._bias.value

    def forward(self, input):
        return input @ self.weights.t() + self.bias

    @property
    def parameters(self):
        return [self._weights, self._bias]

    def emit(self, b, src):
        w_off = b.theta(lambda: self.weights, self.output_dim * self.input_dim)
        b_off = b.theta(lambda: self.bias, self.output_dim)
        return b.op(_lib.GPS_OP_LINEAR, src, self.input_dim, w_off, b_off, n=self.output_dim), self.output_dim

    def symbolic(self, ks):
        out = []
        w = self.weights.detach().cpu().numpy()
        bias = self.bias.detach().cpu().numpy()
        for i in range(self.output_dim):
            tmp = bias[i]
            for j in range(self.input_dim):
                tmp = tmp + ks[j] * w[i, j]
            out.append(tmp)
        return out


class Product(_KernelLayer):
    """Products of consecutive groups of `step` inputs (wrapper.py:132-155)."""

    def __init__(self, input_dim, step, name='Product'):
        super().__init__(input_dim, name=name)
        assert isinstance(step, int) and step > 1, 'step must be number greater than 1'
        assert int(math.fmod(input_dim, step)) == 0, 'input dim must be multiples of step'
        self.step = step
        self.output_dim = input_dim // step

    def forward(self, input):
        return input.reshape(input.shape[0], -1, self.step).prod(-1)

    @property
    def parameters(self):
        return []

    def emit(self, b, src):
        return b.op(_lib.GPS_OP_PRODUCT, src, self.step, n=self.output_dim), self.output_dim

    def symbolic(self, ks):
        return [np.prod(ks[i * self.step:(i + 1) * self.step]) for i in range(int(self.input_dim / self.step))]


class Activation(_KernelLayer):
    """Arbitrary elementwise callable (wrapper.py:158-173): not fusable, evaluated by torch on
    the stacked primitive Grams."""
    fusable = False

    def __init__(self, input_dim, activation_fn, activation_fn_params, name='Activation'):
        super().__init__(input_dim, name=name)
        self.activation_fn = activation_fn
        self.output_dim = input_dim
        self._parameters = activation_fn_params

    def forward(self, input):
        return self.activation_fn(input)

    @property
    def parameters(self):
        return self._parameters

    def symbolic(self, ks):
        return [self.activation_fn(k) for k in ks]
