"""torch.autograd plumbing over the C ABI.

Each Function forwards to one or a few library calls; all O(n^2) / O(n^3) arithmetic --
forward and backward -- runs in the hand-written CUDA kernels.  torch only supplies memory,
the autograd tape and a handful of O(n^2) elementwise masks.
"""
import ctypes
import weakref

import torch

from . import lib as _L
from .lib import TRI_LOWER, TRI_NONE, TRI_UPPER, handle_for, ref, view

F64 = torch.float64


def _prep(t):
    if t.dtype != F64:
        t = t.to(F64)
    if t.dim() == 2 and t.shape[1] > 1 and t.stride(1) != 1:
        t = t.contiguous()
    elif t.dim() == 2 and t.shape[0] > 1 and t.stride(0) < t.shape[1]:
        t = t.contiguous()
    elif t.dim() == 1 and t.numel() > 1 and t.stride(0) != 1:
        t = t.contiguous()
    return t


# ------------------------------------------------------------------------- raw (no autograd)
def gemm_nt(A, B, alpha=1.0, beta=0.0, out=None, a_tri=TRI_NONE, b_tri=TRI_NONE, c_uplo=0):
    """out = alpha * A @ B.T + beta * out  (FP64 tensor-core GEMM)."""
    A, B = _prep(A), _prep(B)
    if A.shape[0] == 0 or B.shape[0] == 0 or A.shape[1] == 0:
        # empty product: nothing to launch (an empty tensor has no device pointer to hand over)
        if out is None:
            return torch.zeros((A.shape[0], B.shape[0]), dtype=F64, device=A.device)
        if A.shape[1] == 0 and out.numel():
            out.mul_(float(beta)) if beta != 0.0 else out.zero_()
        return out
    h = handle_for(A)
    if out is None:
        out = torch.empty((A.shape[0], B.shape[0]), dtype=F64, device=A.device)
        if c_uplo:
            out.zero_()
        beta = 0.0
    va, vb, vc = view(A), view(B), view(out)
    h.check(h.lib.gps_gemm_nt(h.ptr, float(alpha), va.ref, vb.ref, float(beta), vc.ref, a_tri,
                              b_tri, c_uplo))
    return out


def transpose(A):
    A = _prep(A)
    h = handle_for(A)
    out = torch.empty((A.shape[1], A.shape[0]), dtype=F64, device=A.device)
    if A.numel():
        va, vo = view(A), view(out)
        h.check(h.lib.gps_transpose(h.ptr, va.ref, vo.ref))
    return out


_GRAPH_CAPTURE = [False]


def set_graph_capture(flag):
    """While a CUDA graph is being captured (training.GraphedStep) nothing may synchronise the
    stream: the Cholesky skips the device->host read of its `info` flag (a matrix that is not
    positive definite then shows up as NaNs instead of a CholeskyError), and the cached L^-T
    factors are dropped so that the graph records their kernels."""
    _GRAPH_CAPTURE[0] = bool(flag)
    _U_CACHE.clear()


def potrf(K, zero_upper=True, check=True):
    """Returns the lower Cholesky factor of K (K is not modified)."""
    L = _prep(K).clone()
    if L.numel() == 0:
        return L
    h = handle_for(L)
    info = ctypes.c_int(0)
    vl = view(L)
    check = check and not _GRAPH_CAPTURE[0]
    h.check(h.lib.gps_potrf(h.ptr, vl.ref, int(zero_upper), ctypes.byref(info) if check else None))
    return L


def trsm_rlt_(L, B):
    """In place B <- B L^-T."""
    if B.numel() == 0:
        return B
    h = handle_for(B)
    vl, vb = view(_prep(L)), view(B)
    h.check(h.lib.gps_trsm_rlt(h.ptr, vl.ref, vb.ref))
    return B


_U_CACHE = {}


def tri_inv_t(L, share=None):
    """U = L^-T (upper triangular).  Computed once per factor: `share` is the dict that travels
    with a factor produced by `cholesky()` (attribute `_gps_share` of the tensor, also held by
    the autograd nodes that saved the factor), so the forward pass and every adjoint that needs
    U -- Cholesky, triangular solves, L^-T itself -- use one triangular inverse.  Factors that do
    not come from `cholesky()` fall back to a cache keyed by the tensor object."""
    L = _prep(L)
    if L.numel() == 0:
        return L.clone()
    if share is None:
        share = getattr(L, '_gps_share', None)
    if share is not None:
        U = share.get('U')
        if U is not None:
            return U
    key = (L.data_ptr(), L._version, tuple(L.shape))
    hit = _U_CACHE.get(key)
    if share is None and hit is not None and hit[0]() is L:
        return hit[1]
    h = handle_for(L)
    U = torch.empty_like(L, memory_format=torch.contiguous_format)
    vl, vu = view(L), view(U)
    h.check(h.lib.gps_tri_inv_t(h.ptr, vl.ref, vu.ref))
    if share is not None:
        share['U'] = U
    else:
        if len(_U_CACHE) > 8:
            _U_CACHE.clear()
        _U_CACHE[key] = (weakref.ref(L), U)
    return U


def _share_of(L):
    return getattr(L, '_gps_share', None)


def row_sumsq(A):
    A = _prep(A)
    h = handle_for(A)
    out = torch.empty(A.shape[0], dtype=F64, device=A.device)
    if A.shape[0]:
        va, vo = view(A), view(out)
        h.check(h.lib.gps_row_sumsq(h.ptr, 1.0, va.ref, 0.0, vo.ref))
    return out


# ------------------------------------------------------------------------- library-side adjoints
# gps_potri / gps_chol_bwd / gps_trsm_bwd (csrc/adjoint.cu) do in one call what the autograd
# Functions below otherwise compose from gps_tri_inv_t + gps_gemm_nt + transposes: the backward of
# `cholesky` and `trsm_rlt` is one ctypes round trip each.  On by default since it ran green on a
# B200 (profiles/r02_experimental_switches_gpu.txt); False selects the composed formulas.
FUSED_ADJOINTS = [True]


def potri(L):
    """(L L^T)^-1, lower triangle (the strict upper part of the result is zero)."""
    L = _prep(L)
    out = torch.zeros_like(L, memory_format=torch.contiguous_format)
    if L.numel():
        h = handle_for(L)
        vl, vo = view(L), view(out)
        h.check(h.lib.gps_potri(h.ptr, vl.ref, vo.ref))
    return out


def chol_bwd(L, Lbar, U=None):
    """Adjoint of L = chol(A): sym(L^-T Phi(L^T tril(Lbar)) L^-1)."""
    L, Lbar = _prep(L), _prep(Lbar)
    out = torch.empty_like(L, memory_format=torch.contiguous_format)
    if L.numel():
        h = handle_for(L)
        vl, vb, vu, vo = view(L), view(Lbar), view(U), view(out)
        h.check(h.lib.gps_chol_bwd(h.ptr, vl.ref, vb.ref, ref(vu), vo.ref))
    return out


def trsm_bwd(L, X, Xbar, U=None, want_lbar=True):
    """Adjoint of X = B L^-T: (Bbar = Xbar L^-1, Lbar = -tril(Bbar^T X) or None)."""
    L, X, Xbar = _prep(L), _prep(X), _prep(Xbar)
    Bbar = torch.empty_like(X, memory_format=torch.contiguous_format)
    Lbar = torch.empty_like(L, memory_format=torch.contiguous_format) if want_lbar else None
    if X.numel() == 0:
        return Bbar, (torch.zeros_like(L) if want_lbar else None)
    h = handle_for(L)
    vl, vx, vxb, vu, vb, vlb = view(L), view(X), view(Xbar), view(U), view(Bbar), view(Lbar)
    h.check(h.lib.gps_trsm_bwd(h.ptr, vl.ref, vx.ref, vxb.ref, ref(vu), vb.ref, ref(vlb)))
    return Bbar, Lbar


# ------------------------------------------------------------------------- autograd Functions
# Switch (on by default since it ran green on a B200; the CPU tests run both settings): let the
# adjoint of a triangular-aware product skip the zero tiles too.  With C = tri_a(A) tri_b(B)^T:
#   dA = G tri_b(B)      -> NT product with B^T, which is triangular the other way round;
#   dB = tri_b(G^T tri_a(A)) -> only the wanted triangle is computed (lower-output GEMM; an upper
#                           triangle is computed as the lower triangle of the transpose).
# For the SVGP bound with a full q_sqrt (conditionals.py:109-111) this halves two of the four
# N M^2 products of the backward pass.
TRI_AWARE_ADJOINTS = [True]
_FLIP = {TRI_NONE: TRI_NONE, TRI_LOWER: TRI_UPPER, TRI_UPPER: TRI_LOWER}


def _matmul_nt_backward_tri(ctx, A, B, G, a_tri, b_tri):
    dA = dB = None
    if ctx.needs_input_grad[0]:
        dA = gemm_nt(G, transpose(B), b_tri=_FLIP[b_tri])
        if a_tri == TRI_LOWER:
            dA = torch.tril(dA)
        elif a_tri == TRI_UPPER:
            dA = torch.triu(dA)
    if ctx.needs_input_grad[1]:
        if b_tri == TRI_LOWER:
            dB = gemm_nt(transpose(G), transpose(A), b_tri=_FLIP[a_tri], c_uplo=1)
        elif b_tri == TRI_UPPER:
            # upper triangle of G^T A  ==  transpose of the lower triangle of A^T G
            dB = transpose(gemm_nt(transpose(A), transpose(G), a_tri=_FLIP[a_tri], c_uplo=1))
        else:
            dB = gemm_nt(transpose(G), transpose(A), b_tri=_FLIP[a_tri])
    return dA, dB, None, None


class _MatmulNT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, B, a_tri, b_tri):
        A, B = _prep(A), _prep(B)
        ctx.save_for_backward(A, B)
        ctx.tri = (a_tri, b_tri)
        return gemm_nt(A, B, a_tri=a_tri, b_tri=b_tri)

    @staticmethod
    def backward(ctx, G):
        A, B = ctx.saved_tensors
        a_tri, b_tri = ctx.tri
        G = _prep(G)
        dA = dB = None
        if TRI_AWARE_ADJOINTS[0] and (a_tri != TRI_NONE or b_tri != TRI_NONE):
            return _matmul_nt_backward_tri(ctx, A, B, G, a_tri, b_tri)
        if ctx.needs_input_grad[0]:
            # dA = G B   ==  G (B^T)^T
            dA = gemm_nt(G, transpose(B))
            if a_tri == TRI_LOWER:
                dA = torch.tril(dA)
            elif a_tri == TRI_UPPER:
                dA = torch.triu(dA)
        if ctx.needs_input_grad[1]:
            # dB = G^T A  ==  (G^T) (A^T)^T
            dB = gemm_nt(transpose(G), transpose(A))
            if b_tri == TRI_LOWER:
                dB = torch.tril(dB)
            elif b_tri == TRI_UPPER:
                dB = torch.triu(dB)
        return dA, dB, None, None


class _RowSumSq(torch.autograd.Function):
    """sum_k A[i, k]^2 per row (the reference's reduce_sum(square(A), 0) in the transposed
    orientation, conditionals.py:94,118): one pass of gps_row_sumsq, no N x M temporary."""

    @staticmethod
    def forward(ctx, A):
        A = _prep(A)
        ctx.save_for_backward(A)
        return row_sumsq(A)

    @staticmethod
    def backward(ctx, g):
        (A,) = ctx.saved_tensors
        return A * (2.0 * g).unsqueeze(1)


def row_sumsq_ad(A):
    return _RowSumSq.apply(A)


def matmul_nt(A, B, a_tri=TRI_NONE, b_tri=TRI_NONE):
    """A @ B.T; a_tri / b_tri declare A / B triangular (zero tiles are skipped)."""
    return _MatmulNT.apply(A, B, a_tri, b_tri)


def matmul(A, B):
    """A @ B through the NT tensor-core kernel (B is transposed by a device kernel)."""
    return matmul_nt(A, _Transpose.apply(B))


class _Transpose(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A):
        return transpose(A)

    @staticmethod
    def backward(ctx, G):
        return transpose(_prep(G))


def t(A):
    return _Transpose.apply(A)


class _Cholesky(torch.autograd.Function):
    """L = chol(K) (tf.cholesky).  Backward: Kbar = sym(L^-T Phi(L^T Lbar) L^-1) (Murray 2016),
    evaluated with U = L^-T and triangular-aware tensor-core GEMMs."""

    @staticmethod
    def forward(ctx, K, share):
        L = potrf(K)
        ctx.save_for_backward(L)
        ctx.share = share
        return L

    @staticmethod
    def backward(ctx, Lbar):
        (L,) = ctx.saved_tensors
        if FUSED_ADJOINTS[0]:
            return chol_bwd(L, Lbar, tri_inv_t(L, ctx.share)), None
        Lbar = torch.tril(_prep(Lbar))
        U = tri_inv_t(L, ctx.share)
        # P = Phi(L^T Lbar):  (L^T Lbar)[m,n] = sum_k Lt[m,k] Lbar_t[n,k]
        P = gemm_nt(transpose(L), transpose(Lbar), a_tri=TRI_UPPER, b_tri=TRI_UPPER)
        P = torch.tril(P)
        P.diagonal().mul_(0.5)
        # S = U P U^T:  Qt = U P^T ; S = U Qt^T ... with NT products: Qt[m,n] = sum_k U[m,k] P[n,k]
        Qt = gemm_nt(U, P, a_tri=TRI_UPPER, b_tri=TRI_LOWER)
        S = gemm_nt(U, Qt, a_tri=TRI_UPPER)
        return 0.5 * (S + S.t()), None


def cholesky(K):
    share = {}
    L = _Cholesky.apply(K, share)
    L._gps_share = share       # travels with the factor: one L^-T for the forward pass and all adjoints
    return L


class _TrsmRLT(torch.autograd.Function):
    """X = B L^-T  (the row-major form of tf.matrix_triangular_solve(L, B^T, lower=True)^T)."""

    @staticmethod
    def forward(ctx, B, L):
        ctx.share = _share_of(L)
        L = _prep(L)
        X = _prep(B).clone()
        trsm_rlt_(L, X)
        ctx.save_for_backward(X, L)
        return X

    @staticmethod
    def backward(ctx, Xbar):
        X, L = ctx.saved_tensors
        Xbar = _prep(Xbar)
        U = tri_inv_t(L, ctx.share)
        if FUSED_ADJOINTS[0]:
            Bbar, dL = trsm_bwd(L, X, Xbar, U, want_lbar=ctx.needs_input_grad[1])
            return (Bbar if ctx.needs_input_grad[0] else None), dL
        # Bbar = Xbar L^-1 = Xbar U^T
        Bbar = gemm_nt(Xbar, U, b_tri=TRI_UPPER)
        dL = None
        if ctx.needs_input_grad[1]:
            # Lbar = -tril(Bbar^T X)
            dL = gemm_nt(transpose(Bbar), transpose(X), alpha=-1.0, c_uplo=1)
        return (Bbar if ctx.needs_input_grad[0] else None), dL


def trsm_rlt(B, L):
    return _TrsmRLT.apply(B, L)


def solve_lower(L, B):
    """tf.matrix_triangular_solve(L, B, lower=True) for column-layout B [n, m]."""
    return t(trsm_rlt(t(B), L))


def solve_upper_t(L, B):
    """tf.matrix_triangular_solve(tf.transpose(L), B, lower=False) = L^-T B  (conditionals.py:100):
    (L^-T B)^T = B^T L^-1 = B^T U^T with U = L^-T."""
    return t(matmul_nt(t(B), _TriInvT.apply(L), b_tri=TRI_UPPER))


class _TriInvT(torch.autograd.Function):
    """U = L^-T.  Backward: Lbar = -tril((U Ubar^T U)^T)... derived from dU = -U dL^T U."""

    @staticmethod
    def forward(ctx, L):
        share = _share_of(L)
        L = _prep(L)
        U = tri_inv_t(L, share)
        ctx.save_for_backward(U)
        return U

    @staticmethod
    def backward(ctx, Ubar):
        (U,) = ctx.saved_tensors
        Ubar = torch.triu(_prep(Ubar))
        # dU = -U dL^T U  =>  Lbar = -(U^T Ubar U^T)^T = -U Ubar^T U  (lower part)
        T1 = gemm_nt(U, Ubar, a_tri=TRI_UPPER, b_tri=TRI_UPPER)        # U Ubar^T
        Lbar = gemm_nt(T1, transpose(U), alpha=-1.0)                   # (U Ubar^T) U
        return torch.tril(Lbar)


# ------------------------------------------------------------------------- Gram
class KernelProgram(object):
    """A compiled covariance function: ctypes descriptor + the list of constrained parameter
    tensors whose concatenation is `theta`."""

    def __init__(self, desc, pieces, n_theta):
        self.desc = desc
        self.pieces = pieces          # callables returning tensors (constrained values)
        self.n_theta = n_theta

    def theta(self, device):
        vals = []
        for p in self.pieces:
            v = p() if callable(p) else p
            if not isinstance(v, torch.Tensor):
                v = torch.as_tensor(v, dtype=F64, device=device)
            vals.append(v.to(device=device, dtype=F64).reshape(-1))
        return torch.cat(vals)


class _Gram(torch.autograd.Function):
    @staticmethod
    def forward(ctx, theta, X, X2, prog, diag_add):
        X = _prep(X)
        X2 = None if X2 is None else _prep(X2)
        theta = _prep(theta)
        h = handle_for(X)
        n, m = X.shape[0], (X.shape[0] if X2 is None else X2.shape[0])
        K = torch.empty((n, m), dtype=F64, device=X.device)
        if n and m:
            vt, vx, vx2, vk = view(theta), view(X), view(X2), view(K)
            h.check(h.lib.gps_gram_fwd(h.ptr, ctypes.byref(prog.desc), vt.ref, vx.ref, ref(vx2),
                                       float(diag_add), 0, vk.ref))
        ctx.save_for_backward(theta, X, X2 if X2 is not None else torch.empty(0))
        ctx.prog = prog
        ctx.has_x2 = X2 is not None
        return K

    @staticmethod
    def backward(ctx, W):
        theta, X, X2 = ctx.saved_tensors
        X2 = X2 if ctx.has_x2 else None
        prog = ctx.prog
        W = _prep(W)
        h = handle_for(X)
        need_dx = ctx.needs_input_grad[1]
        need_dx2 = ctx.has_x2 and ctx.needs_input_grad[2]
        if W.numel() == 0:
            return (torch.zeros(prog.n_theta, dtype=F64, device=X.device),
                    torch.zeros_like(X) if need_dx else None,
                    torch.zeros_like(X2) if need_dx2 else None, None, None)
        dtheta = torch.empty(prog.n_theta, dtype=F64, device=X.device)
        if need_dx2 and not need_dx:
            # only the second argument wants a gradient (K(Xb, Z) of the sparse models): ONE pass
            # with the roles swapped, K(X, X2)^T = K(X2, X), yields d theta and d X2 together
            dX2 = torch.empty_like(X2)
            Wt = transpose(W)
            vt, vx, vx2, vw, vd, vdx = view(theta), view(X2), view(X), view(Wt), view(dtheta), view(dX2)
            h.check(h.lib.gps_gram_bwd(h.ptr, ctypes.byref(prog.desc), vt.ref, vx.ref, vx2.ref, vw.ref,
                                       vd.ref, vdx.ref))
            return dtheta, None, dX2, None, None
        dX = torch.empty_like(X) if need_dx else None
        if not ctx.has_x2:
            # the library treats W as symmetric in the one-argument case
            W = 0.5 * (W + W.t())
            W = W.contiguous()
        vt, vx, vx2, vw, vd, vdx = view(theta), view(X), view(X2), view(W), view(dtheta), view(dX)
        h.check(h.lib.gps_gram_bwd(h.ptr, ctypes.byref(prog.desc), vt.ref, vx.ref, ref(vx2),
                                   vw.ref, vd.ref, ref(vdx)))
        dX2 = None
        if need_dx2:
            # gradient w.r.t. the second argument: swap roles, transpose the weights
            dX2 = torch.empty_like(X2)
            Wt = transpose(W)
            scratch = torch.empty(prog.n_theta, dtype=F64, device=X.device)
            vwt, vs, vdx2 = view(Wt), view(scratch), view(dX2)
            h.check(h.lib.gps_gram_bwd(h.ptr, ctypes.byref(prog.desc), vt.ref, vx2.ref, vx.ref,
                                       vwt.ref, vs.ref, vdx2.ref))
        return dtheta, dX, dX2, None, None


def gram(prog, X, X2=None, diag_add=0.0):
    theta = prog.theta(X.device)
    return _Gram.apply(theta, X, X2, prog, diag_add)


class _Kdiag(torch.autograd.Function):
    @staticmethod
    def forward(ctx, theta, X, prog):
        X, theta = _prep(X), _prep(theta)
        h = handle_for(X)
        out = torch.empty(X.shape[0], dtype=F64, device=X.device)
        if X.shape[0]:
            vt, vx, vo = view(theta), view(X), view(out)
            h.check(h.lib.gps_kdiag_fwd(h.ptr, ctypes.byref(prog.desc), vt.ref, vx.ref, vo.ref))
        ctx.save_for_backward(theta, X)
        ctx.prog = prog
        return out

    @staticmethod
    def backward(ctx, w):
        theta, X = ctx.saved_tensors
        prog = ctx.prog
        w = _prep(w)
        h = handle_for(X)
        dtheta = torch.empty(prog.n_theta, dtype=F64, device=X.device)
        dX = torch.empty_like(X) if ctx.needs_input_grad[1] else None
        vt, vx, vw, vd, vdx = view(theta), view(X), view(w), view(dtheta), view(dX)
        h.check(h.lib.gps_kdiag_bwd(h.ptr, ctypes.byref(prog.desc), vt.ref, vx.ref, vw.ref, vd.ref,
                                    ref(vdx)))
        return dtheta, dX, None


def kdiag(prog, X):
    return _Kdiag.apply(prog.theta(X.device), X, prog)


# ------------------------------------------------------------------------- fused GPR
class _GprLogLik(torch.autograd.Function):
    """log p(Y) of GPR (models/gpr.py:55-72) with its gradient, one library call."""

    @staticmethod
    def forward(ctx, theta, noise, Yc, X, prog):
        X, Yc, theta = _prep(X), _prep(Yc), _prep(theta)
        h = handle_for(X)
        want_grad = bool(theta.requires_grad or noise.requires_grad or Yc.requires_grad)
        # torch clears requires_grad inside Function.forward; use the ctx flags instead
        want_grad = any(ctx.needs_input_grad[:3])
        scal = torch.zeros(2, dtype=F64, device=X.device)
        dtheta = torch.empty(prog.n_theta, dtype=F64, device=X.device) if want_grad else None
        dY = torch.empty_like(Yc) if (want_grad and ctx.needs_input_grad[2]) else None
        info = ctypes.c_int(0)
        vt, vx, vy, vs, vd, vdy = view(theta), view(X), view(Yc), view(scal), view(dtheta), view(dY)
        h.check(h.lib.gps_gpr_nlml_fwd_bwd(h.ptr, ctypes.byref(prog.desc), vt.ref, vx.ref, vy.ref,
                                           float(noise), int(want_grad), vs.ref, ref(vd), ref(vdy),
                                           ctypes.byref(info)))
        ctx.grads = (dtheta, scal[1] if want_grad else None, dY)
        return -scal[0]

    @staticmethod
    def backward(ctx, g):
        dtheta, dnoise, dY = ctx.grads
        # stored gradients are those of the NEGATIVE log likelihood
        gt = -g * dtheta if ctx.needs_input_grad[0] else None
        gn = -g * dnoise if ctx.needs_input_grad[1] else None
        gy = -g * dY if (ctx.needs_input_grad[2] and dY is not None) else None
        return gt, gn, gy, None, None


class _GprLogLikDist(torch.autograd.Function):
    """log p(Y) of GPR computed by all ranks of a process group together (block-row
    distributed Cholesky / inverse, _backend/dist_gpr.py)."""

    @staticmethod
    def forward(ctx, theta, noise, Yc, X, prog, group, block, lookahead):
        from . import dist_gpr
        X, Yc, theta = _prep(X), _prep(Yc), _prep(theta)
        want_grad = any(ctx.needs_input_grad[:3])
        nlml, dtheta, dnoise, dY = dist_gpr.nlml_and_grad(prog, theta.detach(), float(noise), X, Yc.detach(),
                                                          block=block, group=group, want_grad=want_grad,
                                                          lookahead=lookahead)
        ctx.grads = (dtheta, dnoise, dY)
        return -nlml

    @staticmethod
    def backward(ctx, g):
        dtheta, dnoise, dY = ctx.grads
        gt = -g * dtheta if ctx.needs_input_grad[0] else None
        gn = -g * dnoise if ctx.needs_input_grad[1] else None
        gy = -g * dY if ctx.needs_input_grad[2] else None
        return gt, gn, gy, None, None, None, None, None


def gpr_loglik(prog, X, Yc, noise):
    noise = noise if isinstance(noise, torch.Tensor) else torch.as_tensor(noise, dtype=F64,
                                                                          device=X.device)
    from .. import parallel
    if parallel.active():
        return _GprLogLikDist.apply(prog.theta(X.device), noise.reshape(()), Yc, X, prog,
                                    parallel.group(), parallel.block(), parallel.lookahead())
    return _GprLogLik.apply(prog.theta(X.device), noise.reshape(()), Yc, X, prog)


def gpr_predict(prog, X, Yc, noise, Xnew, full_cov=False):
    X, Yc, Xnew = _prep(X), _prep(Yc), _prep(Xnew)
    theta = _prep(prog.theta(X.device).detach())
    h = handle_for(X)
    ns, r = Xnew.shape[0], Yc.shape[1]
    mean = torch.empty((ns, r), dtype=F64, device=X.device)
    var = torch.empty((ns, ns) if full_cov else (ns,), dtype=F64, device=X.device)
    if ns == 0:
        return mean, var
    info = ctypes.c_int(0)
    vt, vx, vy, vn, vm, vv = view(theta), view(X), view(Yc), view(Xnew), view(mean), view(var)
    h.check(h.lib.gps_gpr_predict(h.ptr, ctypes.byref(prog.desc), vt.ref, vx.ref, vy.ref,
                                  float(noise), vn.ref, int(full_cov), vm.ref, vv.ref,
                                  ctypes.byref(info)))
    return mean, var
