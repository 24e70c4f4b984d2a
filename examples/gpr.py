"""The reference's examples/gpr.py on this package: exact GP regression with a 13-D ARD-RBF
kernel trained by Adam(1e-3) on the NLML, test RMSE / log likelihood printed as it goes.

What changed relative to the reference script (examples/gpr.py:48-80): the TensorFlow session /
placeholder boilerplate is gone (`m.objective` is recomputed on access, like the reference's
eager mode) and `tf.train.AdamOptimizer` became `gpf.training.AdamOptimizer` (same update rule).
The Boston-housing download is replaced by synthetic data of the same shape (455 x 13 train,
51 test): there is no network where this runs.

    python examples/gpr.py [--iters 2000] [--lbfgs]
"""
import argparse
import os.path as osp
import sys

import numpy as np
from scipy import stats

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, osp.join(ROOT, 'gpflow-slim_b200'))

import gpflowSlim as gpf  # noqa: E402


def standardize(data_train, *args):
    """Zero mean / unit standard deviation w.r.t. the training set (examples/gpr.py:12-31)."""
    std = np.std(data_train, 0, keepdims=True)
    std[std == 0] = 1
    mean = np.mean(data_train, 0, keepdims=True)
    return [(d - mean) / std for d in (data_train,) + args] + [mean, std]


def synthetic_housing(seed=1231, n=506, d=13, test_fraction=.1):
    """Stand-in for load_boston_housing (examples/gpr.py:34-40): same shapes, same
    standardisation, a smooth non-linear target with heteroscedastic-free noise."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)) * rng.uniform(0.5, 3.0, size=(1, d))
    w = rng.standard_normal(d) / np.sqrt(d)
    y = 22.0 + 9.0 * np.tanh(x @ w) + 2.0 * np.sin(x[:, 0]) + rng.standard_normal(n)
    perm = rng.permutation(n)
    n_test = int(round(n * test_fraction))
    te, tr = perm[:n_test], perm[n_test:]
    x_t, x_v, _, _ = standardize(x[tr], x[te])
    y_t, y_v, _, train_std = standardize(y[tr], y[te])
    return x_t, y_t, x_v, y_v, float(np.squeeze(train_std))


def evaluate(m, x_test, y_test, std_y_train):
    import torch
    with torch.no_grad():
        mu, cov = m.predict_f(x_test)
    mu, cov = mu.squeeze().cpu().numpy(), cov.squeeze().cpu().numpy()
    rmse = np.mean((mu - y_test) ** 2) ** .5 * std_y_train
    ll = np.mean(np.log(stats.norm.pdf(y_test, loc=mu, scale=cov ** 0.5))) - np.log(std_y_train)
    return rmse, ll


def main(iters=2000, report=100, lbfgs=False, quiet=False):
    x_train, y_train, x_test, y_test, std_y_train = synthetic_housing()
    x_train, y_train = x_train.astype(gpf.settings.float_type), y_train.astype(gpf.settings.float_type)
    x_test, y_test = x_test.astype(gpf.settings.float_type), y_test.astype(gpf.settings.float_type)

    k = gpf.kernels.RBF(13, ARD=True)
    m = gpf.models.GPR(x_train, np.expand_dims(y_train, 1), kern=k)

    if lbfgs:                                   # models/model.py:172 -- the reference's other trainer
        m.optimize()
        obj = float(m.objective.detach())
    else:
        optimizer = gpf.training.AdamOptimizer(1e-3)
        for it in range(iters):
            obj = float(optimizer.minimize(m))
            if it % report == 0 and not quiet:
                rmse, ll = evaluate(m, x_test, y_test, std_y_train)
                print('Iter {}: Loss = {}'.format(it, obj))
                print('test rmse = {}'.format(rmse))
                print('test ll = {}'.format(ll))
    rmse, ll = evaluate(m, x_test, y_test, std_y_train)
    if not quiet:
        print('final: Loss = {}  test rmse = {}  test ll = {}'.format(obj, rmse, ll))
    return obj, rmse, ll


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=2000)
    ap.add_argument('--lbfgs', action='store_true')
    a = ap.parse_args()
    main(iters=a.iters, lbfgs=a.lbfgs)
