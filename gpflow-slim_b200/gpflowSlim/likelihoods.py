"""Likelihoods on the hot path: Gaussian (reference likelihoods.py:30-46, :158-188).
Elementwise [B, 1]-sized host-side maths; the closed-form variational expectation feeds the
SVGP bound."""
import numpy as np
import torch

from . import densities, transforms
from .params import Parameter


class Likelihood(object):
    def __init__(self, name=None):
        self.name = name or type(self).__name__
        self._parameters = []

    @property
    def parameters(self):
        return self._parameters


class Gaussian(Likelihood):
    def __init__(self, var=1.0, min_var=None):
        super().__init__()
        trans = transforms.positive if min_var is None else transforms.Log1pe(min_var)
        self._variance = Parameter(var, transform=trans, name='variance')
        self._parameters = self._parameters + [self._variance]

    @property
    def variance(self):
        return self._variance.value

    def logp(self, F, Y):
        return densities.gaussian(F, Y, self.variance)

    def conditional_mean(self, F):
        return F

    def conditional_variance(self, F):
        return torch.ones_like(F) * self.variance

    def predict_mean_and_var(self, Fmu, Fvar):
        return Fmu, Fvar + self.variance

    def predict_density(self, Fmu, Fvar, Y):
        return densities.gaussian(Fmu, Y, Fvar + self.variance)

    def variational_expectations(self, Fmu, Fvar, Y):
        """likelihoods.py:186-188."""
        return -0.5 * np.log(2 * np.pi) - 0.5 * torch.log(self.variance) \
            - 0.5 * ((Y - Fmu) ** 2 + Fvar) / self.variance
