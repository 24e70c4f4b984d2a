"""TEST INFRASTRUCTURE ONLY -- CPU float64 restatement of the reference's GP hot path.

This is the ORACLE: a plain torch-CPU restatement, op for op, of what the reference's
TensorFlow graph computes on the path  Gram -> Cholesky -> triangular solves -> NLML / ELBO /
collapsed bound -> predictive mean & variance  (gradients by torch autograd, whose Cholesky and
triangular-solve adjoints are mathematically TF's).  Each function cites the reference
file:line it follows (paths relative to /root/reference/gpflowSlim).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it;
the product never does.

Parity status: PINNED to the reference's own Python by tests/test_oracle_golden.py, which
checks this file against tests/golden/*.npz -- vectors produced by running the unmodified
reference over oracle/tf_shim (see gen_golden.py).  What is NOT pinned is TensorFlow's own
Eigen arithmetic inside each op (TF 1.x is not installable here); the reference ships no golden
vectors or known-answer tests of its own for this path.

Kernel specs are nested dicts of CONSTRAINED values (torch tensors, so autograd reaches them):
  {'type': 'rbf'|'matern12'|'matern32'|'matern52'|'exponential',
   'variance': s2, 'lengthscales': l (scalar or [D]), 'active_dims': None|slice|list}
  {'type': 'linear', 'variance': v (scalar or [D]), 'active_dims': ...}
  {'type': 'periodic', 'variance': s2, 'lengthscales': l, 'period': p, 'active_dims': ...}
  {'type': 'sum'|'product', 'children': [spec | scalar tensor, ...]}
  {'type': 'nkn', 'prims': [spec...], 'layers': [('linear', W, b) | ('product', step), ...]}
and, for SURVEY section 8(f) rank 4 (the reference's remaining covariances):
  {'type': 'white'|'constant', 'variance': s2}
  {'type': 'ratquad', 'variance', 'lengthscales', 'alpha', 'active_dims'}
  {'type': 'polynomial', 'variance' (scalar or [D]), 'offset', 'degree', 'active_dims'}
  {'type': 'cosine', 'variance', 'lengthscales', 'weights' [D, 1], 'active_dims'}
  {'type': 'arccosine', 'order', 'variance', 'weight_variances', 'bias_variance', 'active_dims'}
  {'type': 'tps', 'variance', 'active_dims'}
"""
import math

import numpy as np
import torch

F64 = torch.float64
LOG2PI = math.log(2.0 * math.pi)


def T(a):
    return a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a), dtype=F64)


# ------------------------------------------------------------------ transforms (host side)
def softplus_fwd(raw, lower=1e-6):
    """transforms.py:145-146  Log1pe.forward_tensor: softplus(x) + lower."""
    return torch.nn.functional.softplus(T(raw)) + lower


def softplus_inv(y, lower=1e-6):
    """transforms.py:151-178  Log1pe.backward: ys + log(-expm1(-ys)), ys = max(y-lower, eps)."""
    ys = np.maximum(np.asarray(y, dtype=np.float64) - lower, np.finfo(np.float64).eps)
    return ys + np.log(-np.expm1(-ys))


def vec_to_tri(vec, n):
    """misc.py:88-109 / transforms.py:362-365: [K, n(n+1)/2] rows -> [n, n, K] lower
    triangles, filled in numpy.tril_indices (row-major) order."""
    vec = T(vec).reshape(-1, n * (n + 1) // 2)
    r, c = np.tril_indices(n)
    out = torch.zeros(vec.shape[0], n, n, dtype=F64)
    out[:, r, c] = vec
    return out.permute(1, 2, 0)


# ------------------------------------------------------------------ kernels
def _slice(spec, X):
    """kernels.py:217-253  Kernel._slice (slice -> column slice, list -> gather)."""
    ad = spec.get('active_dims')
    if X is None:
        return None
    if ad is None:
        d = spec.get('input_dim')
        return X if d is None else X[:, :d]
    if isinstance(ad, slice):
        return X[:, ad]
    return X[:, list(ad)]


def square_dist(X, X2, ls):
    """kernels.py:408-421  Stationary.square_dist: X/l, -2XX'^T + |x|^2 + |x'|^2, clip >= 0."""
    X = X / ls
    Xs = (X ** 2).sum(1)
    if X2 is None:
        dist = -2.0 * X @ X.T + Xs.reshape(-1, 1) + Xs.reshape(1, -1)
        return torch.clamp(dist, min=0.0)
    X2 = X2 / ls
    X2s = (X2 ** 2).sum(1)
    dist = -2.0 * X @ X2.T + Xs.reshape(-1, 1) + X2s.reshape(1, -1)
    return torch.clamp(dist, min=0.0)


def K(spec, X, X2=None):
    """Gram matrix K(X, X2) of a kernel spec."""
    if not isinstance(spec, dict):                       # scalar constant (kernels.py:1060-1063)
        return T(spec)
    t = spec['type']
    if t in ('sum', 'product'):                          # kernels.py:1071-1084
        vals = [K(c, X, X2) for c in spec['children']]
        out = vals[0]
        for v in vals[1:]:
            out = out + v if t == 'sum' else out * v
        return out
    if t == 'nkn':                                       # neural_kernel_network.py:41-47
        prims = [K(p, X, X2) for p in spec['prims']]
        shp = prims[0].shape
        h = torch.stack([p.reshape(-1) for p in prims], 1)
        return nkn_forward(spec['layers'], h).reshape(shp)
    if t == 'white':                                     # kernels.py:328-338
        if X2 is None:
            return torch.diag_embed(torch.ones(X.shape[0], dtype=F64) * spec['variance'])
        return torch.zeros(X.shape[0], X2.shape[0], dtype=F64)
    if t == 'constant':                                  # kernels.py:341-350
        m = X.shape[0] if X2 is None else X2.shape[0]
        return torch.ones(X.shape[0], m, dtype=F64) * spec['variance']
    Xa, X2a = _slice(spec, X), _slice(spec, X2)
    if t == 'linear':                                    # kernels.py:499-505
        return (Xa * spec['variance']) @ (Xa if X2a is None else X2a).T
    if t == 'polynomial':                                # kernels.py:550-551
        lin = (Xa * spec['variance']) @ (Xa if X2a is None else X2a).T
        return (lin + spec['offset']) ** spec['degree']
    if t == 'cosine':                                    # kernels.py:633-646
        w = spec['weights']
        prod = ((Xa / spec['lengthscales']) @ w).squeeze(-1)
        prod2 = prod if X2a is None else ((X2a / spec['lengthscales']) @ w).squeeze(-1)
        return spec['variance'] * torch.cos(prod[:, None] - prod2[None, :])
    if t == 'arccosine':                                 # kernels.py:740-758
        wv, bv, order = spec['weight_variances'], spec['bias_variance'], spec['order']
        X2b = Xa if X2a is None else X2a
        den = torch.sqrt((wv * Xa ** 2).sum(1) + bv)     # _weighted_product(X) :722-725
        den2 = torch.sqrt((wv * X2b ** 2).sum(1) + bv)
        num = (wv * Xa) @ X2b.T + bv
        cos_theta = num / den[:, None] / den2[None, :]
        jitter = 1e-15
        theta = torch.acos(jitter + (1 - 2 * jitter) * cos_theta)
        return spec['variance'] * (1. / math.pi) * arccos_J(order, theta) \
            * den[:, None] ** order * den2[None, :] ** order
    if t == 'tps':                                       # kernels.py:945-967 (lengthscales unused)
        D = torch.sqrt(square_dist(Xa, X2a, 1.0))
        R_ = 2.0
        return spec['variance'] * (D ** 3 - 1.5 * R_ * D ** 2 + 0.5 * R_ ** 3)
    if t == 'ratquad':                                   # kernels.py:467-471
        d2 = square_dist(Xa, X2a, spec['lengthscales'])
        return spec['variance'] * torch.pow(1.0 + 0.5 * d2 * (1.0 / spec['alpha']), -1.0 * spec['alpha'])
    if t == 'periodic':                                  # kernels.py:806-819
        X2b = Xa if X2a is None else X2a
        r = math.pi * (Xa[:, None, :] - X2b[None, :, :]) / spec['period']
        r = ((torch.sin(r) / spec['lengthscales']) ** 2).sum(2)
        return spec['variance'] * torch.exp(-0.5 * r)
    d2 = square_dist(Xa, X2a, spec['lengthscales'])
    v = spec['variance']
    if t == 'rbf':                                       # kernels.py:436-439
        return v * torch.exp(-d2 / 2.0)
    r = torch.sqrt(d2 + 1e-12)                           # kernels.py:424-426 euclid_dist
    if t == 'exponential':                               # kernels.py:562-566
        return v * torch.exp(-0.5 * r)
    if t == 'matern12':                                  # kernels.py:573-577
        return v * torch.exp(-r)
    if t == 'matern32':                                  # kernels.py:589-594
        return v * (1.0 + math.sqrt(3.0) * r) * torch.exp(-math.sqrt(3.0) * r)
    if t == 'matern52':                                  # kernels.py:605-610
        return v * (1.0 + math.sqrt(5.0) * r + 5.0 / 3.0 * r ** 2) * torch.exp(-math.sqrt(5.0) * r)
    raise ValueError(t)


def Kdiag(spec, X):
    """kernels.py:428-429 (stationary), :507-510 (linear), :803-804 (periodic), :1074-1084."""
    if not isinstance(spec, dict):
        return T(spec)
    t = spec['type']
    if t in ('sum', 'product'):
        vals = [Kdiag(c, X) for c in spec['children']]
        out = vals[0]
        for v in vals[1:]:
            out = out + v if t == 'sum' else out * v
        return out
    if t == 'nkn':                                       # neural_kernel_network.py:35-39
        h = torch.stack([Kdiag(p, X) for p in spec['prims']], 1)
        return nkn_forward(spec['layers'], h).squeeze(-1)
    if t == 'linear':
        Xa = _slice(spec, X)
        return (Xa ** 2 * spec['variance']).sum(1)
    if t == 'polynomial':                                # kernels.py:553-554
        Xa = _slice(spec, X)
        return ((Xa ** 2 * spec['variance']).sum(1) + spec['offset']) ** spec['degree']
    if t == 'arccosine':                                 # kernels.py:760-766
        Xa = _slice(spec, X)
        prod = (spec['weight_variances'] * Xa ** 2).sum(1) + spec['bias_variance']
        return spec['variance'] * (1. / math.pi) * arccos_J(spec['order'], torch.zeros((), dtype=F64)) \
            * prod ** spec['order']
    if t == 'tps':                                       # kernels.py:969-970
        return spec['variance'] * 0.5 * 2.0 ** 3 * torch.ones(X.shape[0], dtype=F64)
    # white / constant (:324-325), stationary incl. ratquad / cosine (:428-429), periodic
    return torch.ones(X.shape[0], dtype=F64) * spec['variance']


def arccos_J(order, theta):
    """kernels.py:727-738: J_0 = pi - t; J_1 = sin t + (pi - t) cos t;
    J_2 = 3 sin t cos t + (pi - t)(1 + 2 cos^2 t)."""
    if order == 0:
        return math.pi - theta
    if order == 1:
        return torch.sin(theta) + (math.pi - theta) * torch.cos(theta)
    if order == 2:
        return 3. * torch.sin(theta) * torch.cos(theta) + (math.pi - theta) * (1. + 2. * torch.cos(theta) ** 2)
    raise ValueError(order)


def nkn_forward(layers, h):
    """neural_kernel_network_wrapper.py:42-47, Linear :114-115, Product :142-145."""
    for layer in layers:
        if layer[0] == 'linear':
            h = h @ layer[1].T + layer[2]
        elif layer[0] == 'product':
            h = h.reshape(h.shape[0], -1, layer[1]).prod(-1)
        else:
            raise ValueError(layer[0])
    return h


# ------------------------------------------------------------------ densities / KL / conditional
def tri_solve(L, B, lower=True):
    """tf.matrix_triangular_solve(L, B, lower) -- reads only the named triangle."""
    return torch.linalg.solve_triangular(torch.tril(L) if lower else torch.triu(L), B,
                                         upper=not lower)


def multivariate_normal(x, mu, L):
    """densities.py:73-95."""
    d = x - mu
    alpha = tri_solve(L, d)
    num_col = 1 if x.dim() == 1 else x.shape[1]
    num_dims = x.shape[0]
    ret = -0.5 * num_dims * num_col * LOG2PI
    ret = ret - num_col * torch.log(torch.diagonal(L)).sum()
    ret = ret - 0.5 * (alpha ** 2).sum()
    return ret


def gauss_kl(q_mu, q_sqrt, Kp=None):
    """kullback_leiblers.py:26-105 (white / non-white, diagonal / full q_sqrt)."""
    white = Kp is None
    if white:
        alpha = q_mu
    else:
        Lp = torch.linalg.cholesky(Kp)
        alpha = tri_solve(Lp, q_mu)
    diag = q_sqrt.dim() == 2
    if diag:
        num_latent = q_sqrt.shape[1]
        NM = q_sqrt.numel()
        Lq = Lq_diag = q_sqrt
    else:
        num_latent = q_sqrt.shape[2]
        NM = q_sqrt.shape[1] * q_sqrt.shape[2]
        Lq = torch.tril(q_sqrt.permute(2, 0, 1))
        Lq_diag = torch.diagonal(Lq, dim1=-2, dim2=-1)
    mahalanobis = (alpha ** 2).sum()
    constant = -float(NM)
    logdet_qcov = torch.log(Lq_diag ** 2).sum()
    if white:
        trace = (Lq ** 2).sum()
    elif diag:
        M = Lp.shape[0]
        Lp_inv = tri_solve(Lp, torch.eye(M, dtype=F64))
        K_inv = tri_solve(Lp.T, Lp_inv, lower=False)
        trace = (torch.diagonal(K_inv)[:, None] * q_sqrt ** 2).sum()
    else:
        LpiLq = tri_solve(Lp.unsqueeze(0).expand(num_latent, -1, -1), Lq)
        trace = (LpiLq ** 2).sum()
    twoKL = mahalanobis + constant - logdet_qcov + trace
    if not white:
        twoKL = twoKL + num_latent * torch.log(torch.diagonal(Lp) ** 2).sum()
    return 0.5 * twoKL


def base_conditional(Kmn, Kmm, Knn, f, full_cov=False, q_sqrt=None, white=False):
    """conditionals.py:81-121."""
    num_func = f.shape[1]
    Lm = torch.linalg.cholesky(Kmm)
    A = tri_solve(Lm, Kmn)
    if full_cov:
        fvar = Knn - A.T @ A
        fvar = fvar.unsqueeze(0).repeat(num_func, 1, 1)
    else:
        fvar = Knn - (A ** 2).sum(0)
        fvar = fvar.unsqueeze(0).repeat(num_func, 1)
    if not white:
        A = tri_solve(Lm.T, A, lower=False)
    fmean = A.T @ f
    if q_sqrt is not None:
        if q_sqrt.dim() == 2:
            LTA = A * q_sqrt.T.unsqueeze(2)
        else:
            L = torch.tril(q_sqrt.permute(2, 0, 1))
            LTA = L.transpose(1, 2) @ A.unsqueeze(0).expand(num_func, -1, -1)
        if full_cov:
            fvar = fvar + LTA.transpose(1, 2) @ LTA
        else:
            fvar = fvar + (LTA ** 2).sum(1)
    fvar = fvar.permute(*reversed(range(fvar.dim())))   # tf.transpose: reverse all axes
    return fmean, fvar


def conditional(spec, Xnew, X, f, full_cov=False, q_sqrt=None, white=False, jitter=1e-6):
    """conditionals.py:25-66 and feature_conditional :70-77 (InducingPoints: features.py:74-81)."""
    Kmm = K(spec, X) + torch.eye(X.shape[0], dtype=F64) * jitter
    Kmn = K(spec, X, Xnew)
    Knn = K(spec, Xnew) if full_cov else Kdiag(spec, Xnew)
    return base_conditional(Kmn, Kmm, Knn, f, full_cov=full_cov, q_sqrt=q_sqrt, white=white)


# ------------------------------------------------------------------ models
def gpr_nlml(spec, X, Y, noise, mean=None):
    """models/gpr.py:55-72 then models/model.py:67-73 (objective = -log p(Y); no priors)."""
    Kxx = K(spec, X) + torch.eye(X.shape[0], dtype=F64) * noise
    L = torch.linalg.cholesky(Kxx)
    m = torch.zeros_like(Y) if mean is None else mean
    return -multivariate_normal(Y, m, L)


def gpr_predict(spec, X, Y, noise, Xnew, full_cov=False):
    """models/gpr.py:118-131 (zero mean function, mean_functions.py:57-59)."""
    Kx = K(spec, X, Xnew)
    Kxx = K(spec, X) + torch.eye(X.shape[0], dtype=F64) * noise
    L = torch.linalg.cholesky(Kxx)
    A = tri_solve(L, Kx)
    V = tri_solve(L, Y)
    fmean = A.T @ V
    if full_cov:
        fvar = K(spec, Xnew) - A.T @ A
        fvar = fvar.unsqueeze(2).repeat(1, 1, Y.shape[1])
    else:
        fvar = Kdiag(spec, Xnew) - (A ** 2).sum(0)
        fvar = fvar.reshape(-1, 1).repeat(1, Y.shape[1])
    return fmean, fvar


def gaussian_var_exp(Fmu, Fvar, Y, var):
    """likelihoods.py:186-188."""
    return -0.5 * LOG2PI - 0.5 * torch.log(var) - 0.5 * ((Y - Fmu) ** 2 + Fvar) / var


def svgp_objective(spec, Xb, Yb, Z, q_mu, q_sqrt, noise, num_data, whiten=True, jitter=1e-6):
    """models/svgp.py:101-130: -(sum var_exp * N/B - KL)."""
    Kuu = None if whiten else K(spec, Z) + torch.eye(Z.shape[0], dtype=F64) * jitter
    KL = gauss_kl(q_mu, q_sqrt, Kuu)
    fmean, fvar = conditional(spec, Xb, Z, q_mu, q_sqrt=q_sqrt, white=whiten, jitter=jitter)
    var_exp = gaussian_var_exp(fmean, fvar, Yb, noise)
    scale = float(num_data) / float(Xb.shape[0])
    return -(var_exp.sum() * scale - KL)


def sgpr_objective(spec, X, Y, Z, noise, jitter=1e-6):
    """models/sgpr.py:121-156."""
    M = Z.shape[0]
    num_data, output_dim = float(Y.shape[0]), float(Y.shape[1])
    err = Y
    kdiag = Kdiag(spec, X)
    Kuf = K(spec, Z, X)
    Kuu = K(spec, Z) + torch.eye(M, dtype=F64) * jitter
    L = torch.linalg.cholesky(Kuu)
    sigma = torch.sqrt(noise)
    A = tri_solve(L, Kuf) / sigma
    AAT = A @ A.T
    B = AAT + torch.eye(M, dtype=F64)
    LB = torch.linalg.cholesky(B)
    Aerr = A @ err
    c = tri_solve(LB, Aerr) / sigma
    bound = -0.5 * num_data * output_dim * LOG2PI
    bound = bound - output_dim * torch.log(torch.diagonal(LB)).sum()
    bound = bound - 0.5 * num_data * output_dim * torch.log(noise)
    bound = bound - 0.5 * (err ** 2).sum() / noise
    bound = bound + 0.5 * (c ** 2).sum()
    bound = bound - 0.5 * output_dim * kdiag.sum() / noise
    bound = bound + 0.5 * output_dim * torch.diagonal(AAT).sum()
    return -bound


def sgpr_predict(spec, X, Y, Z, noise, Xnew, full_cov=False, jitter=1e-6):
    """models/sgpr.py:158-189."""
    M = Z.shape[0]
    Kuf = K(spec, Z, X)
    Kuu = K(spec, Z) + torch.eye(M, dtype=F64) * jitter
    Kus = K(spec, Z, Xnew)
    sigma = torch.sqrt(noise)
    L = torch.linalg.cholesky(Kuu)
    A = tri_solve(L, Kuf) / sigma
    B = A @ A.T + torch.eye(M, dtype=F64)
    LB = torch.linalg.cholesky(B)
    c = tri_solve(LB, A @ Y) / sigma
    tmp1 = tri_solve(L, Kus)
    tmp2 = tri_solve(LB, tmp1)
    mean = tmp2.T @ c
    if full_cov:
        var = K(spec, Xnew) + tmp2.T @ tmp2 - tmp1.T @ tmp1
        var = var.unsqueeze(2).repeat(1, 1, Y.shape[1])
    else:
        var = Kdiag(spec, Xnew) + (tmp2 ** 2).sum(0) - (tmp1 ** 2).sum(0)
        var = var.unsqueeze(1).repeat(1, Y.shape[1])
    return mean, var


def sgpr_upper_bound(spec, X, Y, Z, noise, jitter=1e-6):
    """SGPRUpperMixin.compute_upper_bound, models/sgpr.py:55-82 (Titsias' trace bound)."""
    M = Z.shape[0]
    num_data = float(Y.shape[0])
    kdiag = Kdiag(spec, X)
    Kuu = K(spec, Z) + torch.eye(M, dtype=F64) * jitter
    Kuf = K(spec, Z, X)
    L = torch.linalg.cholesky(Kuu)
    LB = torch.linalg.cholesky(Kuu + noise ** -1.0 * (Kuf @ Kuf.T))
    LinvKuf = tri_solve(L, Kuf)
    c = kdiag.sum() - (LinvKuf ** 2.0).sum()
    corrected_noise = noise + c
    const = -0.5 * num_data * torch.log(2 * math.pi * noise)
    logdet = torch.log(torch.diagonal(L)).sum() - torch.log(torch.diagonal(LB)).sum()
    LC = torch.linalg.cholesky(Kuu + corrected_noise ** -1.0 * (Kuf @ Kuf.T))
    v = tri_solve(LC, corrected_noise ** -1.0 * (Kuf @ Y))
    quad = -0.5 * corrected_noise ** -1.0 * (Y ** 2.0).sum() + 0.5 * (v ** 2.0).sum()
    return const + logdet + quad


def _fitc_common(spec, X, Y, Z, noise, jitter=1e-6):
    """GPRFITC._build_common_terms, models/sgpr.py:227-247."""
    M = Z.shape[0]
    err = Y
    kdiag = Kdiag(spec, X)
    Kuf = K(spec, Z, X)
    Kuu = K(spec, Z) + torch.eye(M, dtype=F64) * jitter
    Luu = torch.linalg.cholesky(Kuu)
    V = tri_solve(Luu, Kuf)
    diagQff = (V ** 2).sum(0)
    nu = kdiag - diagQff + noise
    B = torch.eye(M, dtype=F64) + (V / nu) @ V.T
    L = torch.linalg.cholesky(B)
    beta = err / nu.unsqueeze(1)
    alpha = V @ beta
    gamma = tri_solve(L, alpha)
    return err, nu, Luu, L, alpha, beta, gamma


def gprfitc_objective(spec, X, Y, Z, noise, jitter=1e-6):
    """GPRFITC._build_likelihood, models/sgpr.py:249-291 (negated: Model.objective)."""
    err, nu, Luu, L, alpha, beta, gamma = _fitc_common(spec, X, Y, Z, noise, jitter)
    num_data, num_latent = float(X.shape[0]), float(Y.shape[1])
    mahalanobis = -0.5 * (err ** 2 / nu.unsqueeze(1)).sum() + 0.5 * (gamma ** 2).sum()
    constant = -0.5 * num_data * LOG2PI
    logdet = -0.5 * torch.log(nu).sum() - torch.log(torch.diagonal(L)).sum()
    return -(mahalanobis + (constant + logdet) * num_latent)


def gprfitc_predict(spec, X, Y, Z, noise, Xnew, full_cov=False, jitter=1e-6):
    """GPRFITC._build_predict, models/sgpr.py:293-317."""
    _, _, Luu, L, _, _, gamma = _fitc_common(spec, X, Y, Z, noise, jitter)
    Kus = K(spec, Z, Xnew)
    w = tri_solve(Luu, Kus)
    tmp = tri_solve(L.T, gamma, lower=False)
    mean = w.T @ tmp
    inter = tri_solve(L, w)
    if full_cov:
        var = K(spec, Xnew) - w.T @ w + inter.T @ inter
        var = var.unsqueeze(2).repeat(1, 1, Y.shape[1])
    else:
        var = Kdiag(spec, Xnew) - (w ** 2).sum(0) + (inter ** 2).sum(0)
        var = var.unsqueeze(1).repeat(1, Y.shape[1])
    return mean, var


def tf_adam_step(params, grads, state, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """One step of tf.train.AdamOptimizer (TF 1.x semantics; examples/gpr.py:53-54): epsilon
    sits OUTSIDE the bias-corrected sqrt:  lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    p -= lr_t * m / (sqrt(v) + eps).  (External knowledge of TF 1.x; not in /root/reference.)"""
    state['t'] = state.get('t', 0) + 1
    t = state['t']
    lr_t = lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    out = []
    for i, (p, g) in enumerate(zip(params, grads)):
        m = state.setdefault(('m', i), torch.zeros_like(p))
        v = state.setdefault(('v', i), torch.zeros_like(p))
        m.mul_(b1).add_(g, alpha=1.0 - b1)
        v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
        out.append(p - lr_t * m / (v.sqrt() + eps))
    return out


# ------------------------------------------------------------------ independent cross-check
def gpr_nlml_grad_lapack(X, Y, variance, lengthscales, noise):
    """Independent NumPy/SciPy-LAPACK evaluation of the ARD-RBF GPR objective and its ANALYTIC
    gradient  d/dtheta = 1/2 tr((R K^-1 - beta beta^T) dK/dtheta)  w.r.t. the constrained
    (variance, lengthscales[D], noise).  Used to cross-check the autograd route above (and it
    is the same algebra the CUDA backward uses)."""
    from scipy.linalg import lapack
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    ls = np.asarray(lengthscales, dtype=np.float64) * np.ones(X.shape[1])
    N, R = Y.shape
    Xs = X / ls
    s = (Xs ** 2).sum(1)
    d2 = np.maximum(-2.0 * Xs @ Xs.T + s[:, None] + s[None, :], 0.0)
    Kf = variance * np.exp(-0.5 * d2)
    Kn = Kf + noise * np.eye(N)
    L, info = lapack.dpotrf(Kn, lower=1)
    assert info == 0
    alpha = lapack.dtrtrs(L, Y, lower=1)[0]
    nlml = 0.5 * N * R * LOG2PI + R * np.log(np.diag(L)).sum() + 0.5 * (alpha ** 2).sum()
    Kinv, info = lapack.dpotri(L, lower=1)
    Kinv = np.tril(Kinv) + np.tril(Kinv, -1).T
    beta = lapack.dtrtrs(L, alpha, lower=1, trans=1)[0]
    W = 0.5 * (R * Kinv - beta @ beta.T)
    g_var = (W * Kf).sum() / variance
    g_noise = np.trace(W)
    WK = W * Kf
    g_ls = np.empty(X.shape[1])
    for d in range(X.shape[1]):
        diff2 = (X[:, d:d + 1] - X[:, d:d + 1].T) ** 2
        g_ls[d] = (WK * diff2).sum() / ls[d] ** 3
    return nlml, g_var, g_ls, g_noise
