"""Eager L-BFGS with the step-for-step behaviour of the reference optimiser (LBFGS.py:44-338), the
one `GPModel.optimize` drives (models/model.py:172-195).

Layout for the GPU: all trainable tensors are addressed as ONE flat float64 device vector; the
curvature pairs, the two-loop recursion and the line-search probes are device vector ops, and
only the scalars a branch needs (f, directional derivatives, a handful of dot products) cross
to the host.  An objective evaluation is one fused `gps_gpr_nlml_fwd_bwd` call (or the model's
op sequence) -- the probe `eval_f` costs ONE evaluation, not the two the reference spends
(its `eval_f` evaluates the objective once more only to read the current x, LBFGS.py:79-84).

Reference behaviour that is reproduced on purpose, because iterates would differ otherwise:
  * the curvature pair uses s = d * t with t the INITIAL trial step (min(1, 1/|g|_1) on the
    first iteration, `learningRate` afterwards), not the step the line search accepted
    (LBFGS.py:138, :199-203);
  * `Zoom` receives the directional derivative already reduced to a scalar and multiplies it
    by sum(d) once more (LBFGS.py:275 after :316);
  * line-search evaluations do not count towards `maxEval` (LBFGS.py:222-227);
  * the optimality test is on mean|g| (LBFGS.py:242), the initial one on sum|g| (:112).
"""
import torch


def dot(a, b):
    return (a * b).sum()


def linearize(xs):
    """Concatenate tensors into one flat vector (LBFGS.py:38-42)."""
    return torch.cat([x.reshape(-1) for x in xs])


def _f(v):
    return float(v)


class LBFGS(object):
    """:param opfunc: callable returning (loss, [(grad, var), ...]) -- the contract of
    tfe.implicit_value_and_gradients in the reference; see `model_opfunc`."""

    def __init__(self, opfunc, max_iter=100, lineSearch=True, lineSearchOptions=None,
                 learningRate=1., tolFun=3 * 1e-4, tolX=1e-9, nCorrection=100, verbose=False):
        self.opfunc = opfunc
        self.maxIter = max_iter
        self.maxEval = self.maxIter * 1.25
        self.tolFun = tolFun
        self.tolX = tolX
        self.nCorrection = nCorrection
        self.lineSearch = lineSearch
        self.lineSearchOpts = lineSearchOptions
        self.learningRate = learningRate
        self.isverbose = verbose
        self.history = []
        self.n_evals = 0          # every objective evaluation, line search included

    # ------------------------------------------------------------------ variables
    def get_f_g_x_v(self):
        f, grad_var = self.opfunc()
        self.n_evals += 1
        grads, vars_ = zip(*grad_var)
        grads = [torch.zeros_like(v) if g is None else g for g, v in zip(grads, vars_)]
        self._vars = vars_
        return f.detach(), linearize(grads).detach(), linearize(vars_).detach().clone(), vars_

    def update_vars(self, vars_, x):
        off = 0
        with torch.no_grad():
            for v in vars_:
                n = v.numel()
                v.copy_(x[off:off + n].reshape(v.shape))
                off += n
        assert off == x.shape[0], 'Wrong number of variables'

    def eval_f(self, x):
        """f and g at x; the variables are put back afterwards."""
        vars_ = self._vars
        old_x = linearize(vars_).detach().clone()
        self.update_vars(vars_, x)
        f, g, _, _ = self.get_f_g_x_v()
        self.update_vars(vars_, old_x)
        return f, g

    # ------------------------------------------------------------------ main loop
    def _direction(self, g, pairs, hdiag):
        """Two-loop recursion: -H g for the stored (s, y) pairs."""
        rho = [1.0 / dot(y, s) for s, y in pairs]
        alpha = [None] * len(pairs)
        q = -g
        for i in range(len(pairs) - 1, -1, -1):
            s, y = pairs[i]
            alpha[i] = dot(s, q) * rho[i]
            q = q - alpha[i] * y
        r = q * hdiag
        for i, (s, y) in enumerate(pairs):
            beta = dot(y, r) * rho[i]
            r = r + (alpha[i] - beta) * s
        return r

    def run(self):
        say = print if self.isverbose else (lambda *_: None)
        self.history = []
        f, g, x, vars_ = self.get_f_g_x_v()
        self.history.append((f, g, x, vars_))
        f_hist = [f]
        n_eval = 1
        if _f(g.abs().sum()) <= self.tolFun:
            say('optimality condition below tolFun')
            return x, f_hist

        pairs, hdiag = [], 1.0
        d = t = g_old = None
        for it in range(1, self.maxIter + 1):
            if it == 1:
                d = -g
            else:
                y, s = g - g_old, d * t
                ys = dot(y, s)
                if _f(ys) > 1e-10:
                    if len(pairs) == self.nCorrection:
                        pairs.pop(0)
                    pairs.append((s, y))
                    hdiag = ys / dot(y, y)
                d = self._direction(g, pairs, hdiag)
            g_old, f_old = g, f

            gtd = dot(g, d)
            if _f(gtd) > -self.tolX:
                say('Can not make progress along direction.')
                break
            t = min(1.0, 1.0 / _f(g.abs().sum())) if it == 1 else self.learningRate

            if self.lineSearch:
                step = self.strongwolfe(d, x, f, g)
                self.update_vars(vars_, x + step * d)
            else:
                self.update_vars(vars_, x + t * d)

            if it == self.maxIter:
                break
            f, g, x, vars_ = self.get_f_g_x_v()
            self.history.append((f, g, x, vars_))
            f_hist.append(f)
            n_eval += 1

            if n_eval >= self.maxEval:
                say('max nb of function evals')
                break
            if _f(g.abs().mean()) <= self.tolFun:
                say('optimality condition below tolFun')
                break
            if _f((d * t).abs().sum()) <= self.tolX:
                say('step size below tolX')
                break
            if abs(_f(f) - _f(f_old)) < self.tolX:
                say('function value changing less than tolX')
                break
            if _f(f) - _f(f_old) > 0.5 * abs(_f(f_old)) and it > 10:
                # the objective blew up: step back if the run had been stable, else give up
                prev, before = _f(self.history[-2][0]), _f(self.history[-3][0])
                diff = abs(prev - before)
                if diff < 0.002 * abs(before) or (diff < 0.1 and diff < 0.03 * abs(before)):
                    self.update_vars(vars_, self.history[-2][2])
                    say('Begin to explode, rotating back to previous step')
                    break
                raise ValueError('Very unstable, exit')
        return x, f_hist, n_eval

    # ------------------------------------------------------------------ line search
    def Zoom(self, x0, d, alpha_low, alpha_high, fx0, gx0):
        """Bisection zoom (Nocedal & Wright, algorithm 3.2), at most 6 probes."""
        c1, c2 = 1e-4, 0.9
        fx0 = _f(fx0)
        slope = _f((gx0 * d).sum())
        for trial in range(6):
            mid = 0.5 * (alpha_low + alpha_high)
            f_mid, g_mid = self.eval_f(x0 + mid * d)
            f_mid, dg_mid = _f(f_mid), _f((g_mid * d).sum())
            f_low = _f(self.eval_f(x0 + alpha_low * d)[0])
            if f_mid > fx0 + c1 * mid * slope or f_mid >= f_low:
                alpha_high = mid
            else:
                if abs(dg_mid) <= -c2 * slope:
                    return mid
                if dg_mid * (alpha_high - alpha_low) >= 0:
                    alpha_high = alpha_low
                alpha_low = mid
        return mid

    def strongwolfe(self, d, x0, fx0, gx0):
        """Bracketing phase: trial steps 1, 16.2, 19.24, 19.848 (alpha_max 20, ratio 0.8)."""
        c1, c2 = 1e-4, 0.9
        alpha_max, ratio = 20, 0.8
        slope = (gx0 * d).sum()          # handed on to Zoom as a 0-d tensor, like the reference
        slope_f, fx0_f = _f(slope), _f(fx0)
        prev, f_prev, cur = 0, fx0_f, 1
        for i in range(1, 5):
            f_cur, g_cur = self.eval_f(x0 + cur * d)
            f_cur, dg = _f(f_cur), _f((g_cur * d).sum())
            if f_cur > fx0_f + c1 * cur * slope_f or (i > 1 and f_cur >= f_prev):
                return self.Zoom(x0, d, prev, cur, fx0, slope)
            if abs(dg) <= -c2 * slope_f:
                return cur
            if dg >= 0:
                return self.Zoom(x0, d, cur, prev, fx0, slope)
            prev, f_prev = cur, f_cur
            if i == 4:
                return cur
            cur = cur + (alpha_max - cur) * ratio
        return cur


def model_opfunc(model, var_list=None):
    """(objective, [(grad, tensor), ...]) over the model's trainable tensors -- what
    tfe.implicit_value_and_gradients(lambda: model.objective) yields in the reference
    (models/model.py:174)."""
    def run():
        vs = list(var_list) if var_list is not None else model.trainable_tensors
        obj = model.objective
        gs = torch.autograd.grad(obj, vs, allow_unused=True)
        return obj, list(zip(gs, vs))
    return run
