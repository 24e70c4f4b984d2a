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


class DeviceAdam(object):
    """The same update rule as AdamOptimizer with the step counter and the bias-corrected rate
    kept ON THE DEVICE (0-d tensors), so that a step contains no host-computed scalar and can be
    captured in a CUDA graph."""

    def __init__(self, params, learning_rate=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.params = list(params)
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon
        dev = self.params[0].device
        self.t = torch.zeros((), dtype=torch.float64, device=dev)
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]

    def state(self):
        return [self.t] + self.m + self.v

    def apply_gradients(self, grads):
        with torch.no_grad():
            self.t.add_(1.0)
            lr_t = self.lr * torch.sqrt(1.0 - torch.pow(self.b2, self.t)) / (1.0 - torch.pow(self.b1, self.t))
            for p, g, m, v in zip(self.params, grads, self.m, self.v):
                if g is None:
                    continue
                m.mul_(self.b1).add_(g, alpha=1.0 - self.b1)
                v.mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
                p.sub_(lr_t * m / (v.sqrt() + self.eps))


class GraphedStep(object):
    """One training step -- objective, gradients of
    every trainable tensor, Adam update -- captured ONCE in a CUDA graph and replayed per
    minibatch ("CUDA-graph replay for SVGP").  The SVGP step of BASELINE
    config C4 issues 37 library calls and ~100 torch ops: ~2 ms of host time per step (measured
    against a null library), which replay removes from the critical path.

        step = gpf.training.GraphedStep(model, Xb0, Yb0, learning_rate=1e-3)
        for Xb, Yb in batches:                    # same shapes as Xb0 / Yb0
            objective = step(Xb, Yb)              # 0-d device tensor, valid until the next call

    Constraints: static minibatch shape; nothing in the step may synchronise (the fused one-call
    GPR objective reads its noise as a host scalar and is therefore not capturable -- this is for
    the op-by-op models: SVGP, SGPR, GPMC ...); a non-positive-definite Kuu yields NaNs instead
    of a CholeskyError while graphed."""

    def __init__(self, model, Xb, Yb, learning_rate=1e-3, var_list=None, warmup=2, **adam):
        from ._backend import ops as _ops
        from .misc import to_tensor
        self.model = model
        self.static_X, self.static_Y = to_tensor(Xb).clone(), to_tensor(Yb).clone()
        model.X, model.Y = self.static_X, self.static_Y
        self.params = list(var_list) if var_list is not None else model.trainable_tensors
        self.opt = DeviceAdam(self.params, learning_rate, **adam)
        # warm-up on a side stream: grows the library workspaces, sets kernel attributes, fills
        # the allocator -- then the parameters and the optimiser state are put back
        saved = [t.detach().clone() for t in self.params + self.opt.state()]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        with torch.no_grad():
            for t, s in zip(self.params + self.opt.state(), saved):
                t.copy_(s)
        self.graph = torch.cuda.CUDAGraph()
        _ops.set_graph_capture(True)
        try:
            # capture on the stream the warm-up ran on: per-stream library state (the split-K
            # scratch of the GEMM) then exists already -- no allocation may happen while capturing
            with torch.cuda.graph(self.graph, stream=side):
                self.static_objective = self._step()
        finally:
            _ops.set_graph_capture(False)

    def _step(self):
        obj = self.model.objective
        grads = torch.autograd.grad(obj, self.params, allow_unused=True)
        self.opt.apply_gradients(grads)
        return obj.detach()

    def __call__(self, Xb, Yb):
        self.static_X.copy_(Xb)
        self.static_Y.copy_(Yb)
        self.graph.replay()
        return self.static_objective

