"""TEST DOUBLE (never shipped): torch-CPU stand-ins for `gpflowSlim._backend.ops`, so that the
package's Python HOST LOGIC -- kernel-expression compilation, the composed kernels, models,
conditionals, likelihood glue, optimisers -- runs on a machine without a GPU and can be held to
the golden vectors there.  The product itself has no CPU path: outside this fixture every one
of these entry points raises without libgpslim_b200.so and an sm_100 device.

The Gram arithmetic comes from the oracle (`oracle/ref_torch.py`): a compiled
`gps_kernel_desc` is walked primitive by primitive / op by op, which also checks that the
descriptor the host code emits means what the CUDA interpreter (csrc/gram.cu) takes it to mean.
"""
import torch

from oracle import ref_torch as R

F64 = torch.float64
TRI_NONE, TRI_LOWER, TRI_UPPER = 0, 1, 2

_STAT = {0: 'rbf', 1: 'exponential', 2: 'matern12', 3: 'matern32', 4: 'matern52'}
OP_CONST, OP_ADD, OP_MUL, OP_COPY, OP_LINEAR, OP_PRODUCT = range(6)


def _tri(A, mode):
    if mode == TRI_LOWER:
        return torch.tril(A)
    if mode == TRI_UPPER:
        return torch.triu(A)
    return A


# --------------------------------------------------------------------------- descriptor walk
def _prim_spec(pr, theta):
    dims = [int(pr.dims[k]) for k in range(pr.ndims)]
    off = pr.theta_off
    if pr.type in _STAT:
        nls = pr.ndims if pr.ard else 1
        ls = theta[off + 1: off + 1 + nls]
        return dict(type=_STAT[pr.type], variance=theta[off], lengthscales=ls if pr.ard else ls[0],
                    active_dims=dims)
    if pr.type == 5:
        v = theta[off: off + pr.ndims] if pr.ard else theta[off]
        return dict(type='linear', variance=v, active_dims=dims)
    if pr.type == 6:
        return dict(type='periodic', variance=theta[off], lengthscales=theta[off + 1],
                    period=theta[off + 2], active_dims=dims)
    raise ValueError('primitive type %d' % pr.type)


def _run_program(desc, theta, prim_vals):
    slots = list(prim_vals)
    P = desc.n_prims
    assert len(slots) == P
    like = prim_vals[0]
    for q in range(desc.n_ops):
        o = desc.ops[q]
        assert o.dst == len(slots), 'ops must append slots in order'
        if o.op == OP_CONST:
            slots.append(theta[o.a] * torch.ones_like(like))
        elif o.op == OP_ADD:
            slots.append(slots[o.a] + slots[o.b])
        elif o.op == OP_MUL:
            slots.append(slots[o.a] * slots[o.b])
        elif o.op == OP_COPY:
            slots.append(slots[o.a])
        elif o.op == OP_LINEAR:
            W = theta[o.c: o.c + o.n * o.b].reshape(o.n, o.b)
            bias = theta[o.d: o.d + o.n]
            inp = torch.stack(slots[o.a: o.a + o.b], -1)
            out = inp @ W.t() + bias
            slots.extend(out[..., r] for r in range(o.n))
        elif o.op == OP_PRODUCT:
            for g in range(o.n):
                v = slots[o.a + g * o.b]
                for c in range(1, o.b):
                    v = v * slots[o.a + g * o.b + c]
                slots.append(v)
        else:
            raise ValueError('op %d' % o.op)
    return slots[desc.out_slot]


def _gram_desc(prog, theta, X, X2):
    d = prog.desc
    assert theta.numel() == d.n_theta == prog.n_theta
    vals = [R.K(_prim_spec(d.prims[i], theta), X, X2) for i in range(d.n_prims)]
    return _run_program(d, theta, vals)


def _kdiag_desc(prog, theta, X):
    d = prog.desc
    vals = [R.Kdiag(_prim_spec(d.prims[i], theta), X) for i in range(d.n_prims)]
    return _run_program(d, theta, vals)


# --------------------------------------------------------------------------- the stand-ins
def gemm_nt(A, B, alpha=1.0, beta=0.0, out=None, a_tri=TRI_NONE, b_tri=TRI_NONE, c_uplo=0):
    res = alpha * (_tri(A, a_tri) @ _tri(B, b_tri).t())
    if out is None:
        return torch.tril(res) if c_uplo else res
    if c_uplo:
        out.copy_(torch.tril(res + beta * out) + torch.triu(out, 1))
    else:
        out.copy_(res + beta * out)
    return out


def transpose(A):
    return A.t().contiguous()


def potrf(K, zero_upper=True, check=True):
    from gpflowSlim._backend.lib import CholeskyError
    S = torch.tril(K) + torch.tril(K, -1).t()
    L, info = torch.linalg.cholesky_ex(S)
    if int(info) != 0:
        raise CholeskyError('leading minor of order %d is not positive definite' % int(info))
    return L if zero_upper else L + torch.triu(K, 1)


def trsm_rlt_(L, B):
    B.copy_(torch.linalg.solve_triangular(torch.tril(L), B.t(), upper=False).t())
    return B


def tri_inv_t(L, share=None):
    n = L.shape[0]
    return torch.linalg.solve_triangular(torch.tril(L), torch.eye(n, dtype=F64), upper=False).t().contiguous()


def row_sumsq(A):
    return (A ** 2).sum(1)


def row_sumsq_ad(A):
    return (A ** 2).sum(1)


def matmul_nt(A, B, a_tri=TRI_NONE, b_tri=TRI_NONE):
    return _tri(A, a_tri) @ _tri(B, b_tri).t()


def matmul(A, B):
    return A @ B


def t(A):
    return A.t()


def cholesky(K):
    from gpflowSlim._backend.lib import CholeskyError
    L, info = torch.linalg.cholesky_ex(K)
    if int(info) != 0:
        raise CholeskyError('leading minor of order %d is not positive definite' % int(info))
    return L


def trsm_rlt(B, L):
    return torch.linalg.solve_triangular(torch.tril(L), B.t(), upper=False).t()


def solve_lower(L, B):
    return torch.linalg.solve_triangular(torch.tril(L), B, upper=False)


def solve_upper_t(L, B):
    return torch.linalg.solve_triangular(torch.tril(L).t(), B, upper=True)


class _TriInvT(object):
    @staticmethod
    def apply(L):
        n = L.shape[0]
        return torch.linalg.solve_triangular(torch.tril(L), torch.eye(n, dtype=F64), upper=False).t()


def gram(prog, X, X2=None, diag_add=0.0):
    K = _gram_desc(prog, prog.theta(X.device), X, X2)
    if diag_add:
        K = K + diag_add * torch.eye(K.shape[0], dtype=F64)
    return K


def kdiag(prog, X):
    return _kdiag_desc(prog, prog.theta(X.device), X)


def gpr_loglik(prog, X, Yc, noise):
    K = gram(prog, X) + noise * torch.eye(X.shape[0], dtype=F64)
    return R.multivariate_normal(Yc, torch.zeros_like(Yc), cholesky(K))


def gpr_predict(prog, X, Yc, noise, Xnew, full_cov=False):
    with torch.no_grad():
        L = cholesky(gram(prog, X) + noise * torch.eye(X.shape[0], dtype=F64))
        A = solve_lower(L, gram(prog, X, Xnew))
        V = solve_lower(L, Yc)
        mean = A.t() @ V
        if full_cov:
            return mean, gram(prog, Xnew) - A.t() @ A
        return mean, kdiag(prog, Xnew) - (A ** 2).sum(0)


NAMES = ['gemm_nt', 'transpose', 'potrf', 'trsm_rlt_', 'tri_inv_t', 'row_sumsq', 'row_sumsq_ad', 'matmul_nt',
         'matmul', 't', 'cholesky', 'trsm_rlt', 'solve_lower', 'solve_upper_t', '_TriInvT', 'gram',
         'kdiag', 'gpr_loglik', 'gpr_predict']


# 'raw' level: only the entry points that touch the library directly are replaced; the package's
# own torch.autograd Functions (_MatmulNT, _Cholesky, _TrsmRLT, _TriInvT, _Transpose: the
# HAND-WRITTEN adjoints of ops.py) stay in place and run on top of the CPU gemm / transpose /
# potrf / trsm / inverse stand-ins -- so their backward formulas, and the triangular-structure
# flags they pass, are exercised on the CPU as well.
RAW_NAMES = ['gemm_nt', 'transpose', 'potrf', 'trsm_rlt_', 'tri_inv_t', 'row_sumsq', 'gram', 'kdiag',
             'gpr_loglik', 'gpr_predict']


def install(monkeypatch, level='ops'):
    """Swap the stand-ins into gpflowSlim._backend.ops for the duration of one test."""
    from gpflowSlim._backend import ops
    g = globals()
    for name in (NAMES if level == 'ops' else RAW_NAMES):
        assert hasattr(ops, name), name
        monkeypatch.setattr(ops, name, g[name])
    if level != 'ops':
        ops._U_CACHE.clear()
    return ops
