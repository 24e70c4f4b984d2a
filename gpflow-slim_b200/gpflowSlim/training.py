"""Optimiser with the reference scripts' semantics: tf.train.AdamOptimizer (examples/gpr.py:53,
examples/svgp.py:161).  TF-1.x Adam puts epsilon OUTSIDE the bias-corrected square root
(lr_t = lr sqrt(1-b2^t)/(1-b1^t); p -= lr_t m / (sqrt(v) + eps)), which torch.optim.Adam does
not reproduce bit for bit -- so config-1 step parity needs this one."""
import math

import torch


class AdamOptimizer(object):
    def __init__(self, learning_rate=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon
        self.t = 0
        self.m, self.v = {}, {}

    def minimize(self, model_or_objective, var_list=None):
        """One step on `model.objective` (or on a callable returning the objective)."""
        if var_list is None:
            var_list = model_or_objective.trainable_tensors
        obj = model_or_objective.objective if hasattr(model_or_objective, 'objective') \
            else model_or_objective()
        grads = torch.autograd.grad(obj, var_list, allow_unused=True)
        self.apply_gradients(zip(grads, var_list))
        return obj.detach()

    def apply_gradients(self, grads_and_vars):
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        with torch.no_grad():
            for g, p in grads_and_vars:
                if g is None:
                    continue
                k = id(p)
                if k not in self.m:
                    self.m[k], self.v[k] = torch.zeros_like(p), torch.zeros_like(p)
                self.m[k].mul_(self.b1).add_(g, alpha=1.0 - self.b1)
                self.v[k].mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
                p.sub_(lr_t * self.m[k] / (self.v[k].sqrt() + self.eps))
