"""GP conditionals (reference conditionals.py:25-121).

Everything is kept in the TRANSPOSED (row-major friendly) orientation: with
At = Kmn^T Lm^-T = (Lm^-1 Kmn)^T of shape [N, M], every product below is an `A B^T` with both
operands K-contiguous -- the one form the FP64 tensor-core GEMM implements -- and the
column reductions of the reference (`reduce_sum(square(A), 0)`, :94, :118) become contiguous
row reductions."""
import torch

from ._backend import ops as _ops
from ._backend.lib import TRI_LOWER, TRI_UPPER
from ._settings import SETTINGS as settings
from .misc import to_tensor


def conditional(Xnew, X, kern, f, *, full_cov=False, q_sqrt=None, white=False):
    """conditionals.py:25-66."""
    Xnew, X = to_tensor(Xnew), to_tensor(X)
    Kmm = kern.K_jittered(X, settings.numerics.jitter_level)
    Knm = kern.K(Xnew, X)                 # Kmn^T, evaluated directly in the orientation used below
    Knn = kern.K(Xnew) if full_cov else kern.Kdiag(Xnew)
    return _base_conditional_t(Knm, Kmm, Knn, f, full_cov=full_cov, q_sqrt=q_sqrt, white=white)


def feature_conditional(Xnew, feat, kern, f, *, full_cov=False, q_sqrt=None, white=False):
    """conditionals.py:70-77."""
    Xnew = to_tensor(Xnew)
    Kmm = feat.Kuu(kern, jitter=settings.numerics.jitter_level)
    Knm = feat.Kfu(kern, Xnew)            # Kuf^T [N, M]
    Knn = kern.K(Xnew) if full_cov else kern.Kdiag(Xnew)
    return _base_conditional_t(Knm, Kmm, Knn, f, full_cov=full_cov, q_sqrt=q_sqrt, white=white)


def base_conditional(Kmn, Kmm, Knn, f, *, full_cov=False, q_sqrt=None, white=False):
    """conditionals.py:81-121.  Kmn [M, N], Kmm [M, M], Knn [N] or [N, N], f [M, K],
    q_sqrt None | [M, K] | [M, M, K]  ->  fmean [N, K], fvar [N, K] or [N, N, K]."""
    return _base_conditional_t(_ops.t(to_tensor(Kmn)), Kmm, Knn, f, full_cov=full_cov, q_sqrt=q_sqrt,
                               white=white)


def _base_conditional_t(Knm, Kmm, Knn, f, *, full_cov=False, q_sqrt=None, white=False):
    """base_conditional with the cross-covariance handed over as Knm = Kmn^T [N, M] (what
    `conditional` / `feature_conditional` evaluate directly: no 8 N M byte transpose)."""
    Knm, Kmm, Knn, f = to_tensor(Knm), to_tensor(Kmm), to_tensor(Knn), to_tensor(f)
    num_func = f.shape[1]
    Lm = _ops.cholesky(Kmm)                                   # :84
    At = _ops.trsm_rlt(Knm, Lm)                               # :87   At = A^T, [N, M]
    if full_cov:
        fvar = Knn - _ops.matmul_nt(At, At)                   # :90   A^T A
        fvar = fvar.unsqueeze(0).expand(num_func, -1, -1)
    else:
        fvar = Knn - _ops.row_sumsq_ad(At)                    # :94
        fvar = fvar.unsqueeze(0).expand(num_func, -1)
    if not white:
        # A <- Lm^-T A  (:99-100);  At <- At Lm^-1 = At U^T with U = Lm^-T
        At = _ops.matmul_nt(At, _ops._TriInvT.apply(Lm), b_tri=TRI_UPPER)
    fmean = _ops.matmul_nt(At, _ops.t(f))                     # :103  A^T f
    if q_sqrt is not None:
        q_sqrt = to_tensor(q_sqrt)
        if q_sqrt.dim() == 2:
            LTAt = At.unsqueeze(0) * q_sqrt.t().unsqueeze(1)                  # :106  [K, N, M]
            if full_cov:
                add = torch.stack([_ops.matmul_nt(LTAt[k], LTAt[k]) for k in range(num_func)])
            else:
                add = (LTAt ** 2).sum(2)
        elif q_sqrt.dim() == 3:
            # LTA_k = L_k^T A  (:109-111)  ->  LTA_k^T = At L_k = At (L_k^T)^T, L_k^T upper
            adds = []
            for k in range(num_func):
                Lkt = _ops.t(torch.tril(q_sqrt[:, :, k]))
                LTAt = _ops.matmul_nt(At, Lkt, b_tri=TRI_UPPER)               # [N, M]
                adds.append(_ops.matmul_nt(LTAt, LTAt) if full_cov else _ops.row_sumsq_ad(LTAt))
            add = torch.stack(adds)
        else:
            raise ValueError('Bad dimension for q_sqrt: %s' % str(q_sqrt.dim()))
        fvar = fvar + add                                     # :115-118
    fvar = fvar.permute(*reversed(range(fvar.dim())))         # :119  N x K or N x N x K
    return fmean, fvar
