"""The reference's OWN two tests on the hot path, restated against this
package's API so they run on the CUDA kernels (tests/test_gpu_zz_widened.py) and on the CPU
test double (tests/test_reference_own_cpu.py):

  * gpflowSlim/densities.py:159-174  Test_multivariate_normal_feature.test_logp --
    log N(x | mu, C C^T + var I) through the N x N Cholesky equals the Woodbury form;
  * gpflowSlim/models/gpr.py:135-203  TestPredict.test_predict -- GPR prediction through the
    feature expansion equals the standard Cholesky predictor on K = C C^T.

Same shapes and constants as the reference (10 x 5 features, var = 2; Nx, Nn, d = 20, 10, 5,
variance = 2); the reference runs them in float32 with assertAllClose defaults / 1e-4.  Here
they run in float64: the predictors must agree to 1e-9, the two densities to 1e-6 (the Woodbury
form adds the 1e-6 jitter to diag(L) inside its log-determinant, densities.py:111)."""
import numpy as np
import torch


class _FixedFeatures(object):
    """A 'kernel' whose feature map is a lookup of fixed random features (the reference's test
    draws feat / feat_new directly, gpr.py:146-147)."""

    def __init__(self, table):
        self.table = table          # {n_rows: features}
        self.parameters = []

    def features(self, X):
        return self.table[X.shape[0]]


def mvn_feature_vs_cholesky(gpf, conv, seed=0):
    rng = np.random.default_rng(seed)
    C = conv(rng.standard_normal((10, 5)))
    var = conv(np.array(2.0))
    mu = conv(np.zeros((10, 1)))
    x = conv(rng.standard_normal((10, 1)))
    from gpflowSlim._backend import ops
    CCt_I = ops.matmul_nt(C, C) + var * torch.eye(10, dtype=C.dtype, device=C.device)
    L = ops.cholesky(CCt_I)
    logp1 = gpf.densities.multivariate_normal(x, mu, L)
    logp2 = gpf.densities.multivariate_normal_feature(x, mu, C, var)
    return float(logp1), float(logp2)


def predict_feature_vs_standard(gpf, conv, seed=1):
    rng = np.random.default_rng(seed)
    Nx, Nn, d = 20, 10, 5
    variance = 2.
    Y = conv(rng.standard_normal((Nx, 1)))
    X = conv(rng.standard_normal((Nx, d)))
    Xnew = conv(rng.standard_normal((Nn, d)))
    feat = conv(rng.standard_normal((Nx, d)))
    feat_new = conv(rng.standard_normal((Nn, d)))

    class _Mean(gpf.mean_functions.MeanFunction):
        def __init__(self, table):
            super().__init__()
            self.table = table

        def __call__(self, Xq):
            return self.table[Xq.shape[0]]
    mean = _Mean({Nx: conv(rng.standard_normal((Nx, 1))), Nn: conv(rng.standard_normal((Nn, 1)))})

    # feature path: GPR sees kern.features
    m_feat = gpf.models.GPR(X, Y, kern=_FixedFeatures({Nx: feat, Nn: feat_new}), mean_function=mean,
                            obs_var=variance)
    fmean, fvar_diag = m_feat._build_predict(Xnew)
    _, fvar_full = m_feat._build_predict(Xnew, full_cov=True)

    # standard path on K = feat feat^T, written out as in the reference test (gpr.py:184-198)
    from gpflowSlim._backend import ops
    var = m_feat.likelihood.variance
    Kx = ops.matmul_nt(feat, feat_new)
    K = ops.matmul_nt(feat, feat) + torch.eye(Nx, dtype=feat.dtype, device=feat.device) * var
    L = ops.cholesky(K)
    A = ops.solve_lower(L, Kx)
    V = ops.solve_lower(L, Y - mean(X))
    s_fmean = ops.matmul_nt(ops.t(A), ops.t(V)) + mean(Xnew)
    s_full = ops.matmul_nt(feat_new, feat_new) - ops.matmul_nt(ops.t(A), ops.t(A))
    s_diag = torch.diagonal(ops.matmul_nt(feat_new, feat_new)) - (A ** 2).sum(0)
    out = lambda t: t.detach().cpu().numpy()
    return (out(fmean), out(s_fmean)), (out(fvar_full[:, :, 0]), out(s_full)), \
        (out(fvar_diag[:, 0]), out(s_diag))
