"""TEST DOUBLE (never shipped): torch-CPU stand-in for gpflowSlim._backend.dist_gpr.CudaBackend,
so the host logic of the distributed GPR path (ownership maps, panel packing / unpacking, masks,
prefix tables, collectives) runs under gloo on a machine without a GPU.  The Gram arithmetic comes
from the oracle (`spec_fn(theta)` -> oracle kernel spec)."""
import torch

from oracle import ref_torch as R

F64 = torch.float64


class CpuBackend(object):
    def __init__(self, spec_fn):
        self.spec_fn = spec_fn

    # look-ahead "streams": sequential execution in issue order is one valid schedule of the
    # dependency graph, so the look-ahead ORDER of operations is what gets tested here
    def streams(self):
        return 'main', 'chain', 'tb', 'gather', 'narrow'

    def on(self, stream):
        import contextlib
        return contextlib.nullcontext()

    def record(self, stream):
        return None

    def wait(self, stream, event):
        pass

    def empty(self, *shape):
        return torch.full(tuple(shape), float('nan'), dtype=F64)   # poison: unwritten reads show up

    def zeros(self, *shape):
        return torch.zeros(*shape, dtype=F64)

    def gram_rows(self, prog, theta, Xr, Xc, out):
        out.copy_(R.K(self.spec_fn(theta), Xr, Xc))

    def potrf_(self, A):
        L = torch.linalg.cholesky(torch.tril(A) + torch.tril(A, -1).T)
        A.copy_(torch.tril(L) + torch.triu(A, 1))          # strict upper untouched, like the library

    def trsm_rlt_(self, Lm, B):
        B.copy_(torch.linalg.solve_triangular(torch.tril(Lm), B.T, upper=False).T)

    def copy_(self, dst, src):
        dst.copy_(src)

    def zero_(self, t):
        t.zero_()

    def unpack_rows_(self, dst, src, index):
        dst.copy_(src.index_select(0, index))

    def syrk_lower_(self, X, D):
        D.sub_(torch.tril(X @ X.T))

    def gemm_rowmap_(self, A, B, C, rowlim, coff, flops=-1.0):
        cols = torch.arange(C.shape[1]) + coff
        mask = cols[None, :] <= rowlim[:, None]
        upd = A @ B.T
        C.sub_(torch.where(mask, upd, torch.zeros_like(upd)))

    def transpose(self, A):
        return A.T.contiguous()

    def transpose_into(self, A, out):
        out.copy_(A.T)

    def trsm_rlt_prefix_(self, Lm, B, row_start):
        # the real kernel never reads what lies left of row_start: poison it to prove that
        n = Lm.shape[0]
        cols = torch.arange(n)[None, :]
        left = cols < torch.as_tensor(row_start)[:, None]
        assert bool((B[left] == 0).all())
        X = torch.linalg.solve_triangular(torch.tril(Lm), B.T, upper=False).T
        B.copy_(torch.where(left, torch.zeros_like(X), X))

    def trsm_rln_prefix_(self, Lm, Lt, B, row_start):
        assert torch.equal(torch.triu(Lt), torch.tril(Lm).T)
        n = Lm.shape[0]
        cols = torch.arange(n)[None, :]
        left = cols < torch.as_tensor(row_start)[:, None]
        X = torch.linalg.solve_triangular(torch.tril(Lm), B, upper=False, left=False)
        # left of row_start the result is undefined in the real kernel: poison it
        B.copy_(torch.where(left, torch.full_like(X, float('nan')), X))

    def weight_rows_(self, W, grow, beta, block):
        n = W.shape[1]
        cols = torch.arange(n)
        c0 = (grow // block * block)[:, None]
        bb = beta[:, grow].T @ beta                        # [m, N]
        v = 0.5 * (beta.shape[0] * W - bb)
        v = torch.where(cols[None, :] >= c0 + block, 2.0 * v, v)
        W.copy_(torch.where(cols[None, :] >= c0, v, torch.zeros_like(v)))   # drops the poison

    def gram_bwd(self, prog, theta, Xr, Xc, W):
        th = theta.detach().clone().requires_grad_(True)
        K = R.K(self.spec_fn(th), Xr, Xc)
        (g,) = torch.autograd.grad((K * W).sum(), th)
        return g

    def matmul_nt(self, A, B):
        return A @ B.T

    def row_sumsq(self, A):
        return (A ** 2).sum(1)

    def sum_log_diag(self, Lm):
        return torch.log(torch.diagonal(Lm)).sum()
