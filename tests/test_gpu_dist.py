"""GPU tests of the multi-GPU building blocks (through the C ABI) and of the distributed GPR /
SVGP paths: kernels against torch fp64, the distributed path at world size 1 against the
fused single-GPU path and the oracle, and -- when the box has >= 2 GPUs -- a 2-rank NCCL run
(tools/dist_check.py under torchrun) against the same values.  Tolerance 1e-8 relative
(north star) for objectives / gradients, 1e-12 for single kernels."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from util import assert_close, conv, dev

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _be():
    from gpflowSlim._backend.dist_gpr import CudaBackend
    return CudaBackend(dev())


def _spd(n, seed=0):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n + 3))
    return A @ A.T / (n + 3) + 0.5 * np.eye(n)


@pytest.mark.parametrize('m,n,k,coff', [(300, 517, 128, 0), (1000, 900, 256, 130), (129, 64, 512, 5),
                                        (2600, 2300, 128, 130)])
def test_gemm_rowmap_masks_by_global_row(m, n, k, coff):
    rng = np.random.default_rng(m + n)
    A, B, C = rng.standard_normal((m, k)), rng.standard_normal((n, k)), rng.standard_normal((m, n))
    lim = np.sort(rng.integers(0, n + coff + 40, size=m))
    lim[-3:] = 1 << 60                                     # ride-along rows: never masked
    ref = C.copy()
    mask = (np.arange(n)[None, :] + coff) <= lim[:, None]
    ref[mask] -= (A @ B.T)[mask]
    Cd = conv(C)
    _be().gemm_rowmap_(conv(A), conv(B), Cd, torch.tensor(lim, device=dev()), coff)
    assert_close(Cd, ref, 1e-13, 'rowmap gemm')


@pytest.mark.parametrize('leaf', [128, 256, 512, 1024])
@pytest.mark.parametrize('n,bs', [(384, 128), (1000, 256), (1500, 384), (4000, 512)])
def test_prefix_triangular_solves(n, bs, leaf):
    """Rows of U = L^-T and of K^-1 for a subset of block rows, against dense inverses; with the
    aligned diagonal blocks of size `leaf` solved by one product with their explicit inverse
    (option trsm_leaf; 128 = strips only, 512 = default)."""
    from gpflowSlim._backend.lib import handle_for
    h = handle_for(dev())
    h.set_option('trsm_leaf', leaf)
    try:
        _prefix_case(n, bs)
    finally:
        h.set_option('trsm_leaf', 512)


def _prefix_case(n, bs):
    S = _spd(n, seed=n)
    L = np.linalg.cholesky(S)
    U, Kinv = np.linalg.inv(L).T, np.linalg.inv(S)
    nblk = (n + bs - 1) // bs
    mine = [b for b in range(nblk) if b % 2 == 0]
    rows = np.concatenate([np.arange(b * bs, min(n, (b + 1) * bs)) for b in mine])
    B = np.zeros((len(rows), n))
    B[np.arange(len(rows)), rows] = 1.0
    act = rows // bs * bs                                  # first column of every row
    be = _be()
    Ld = conv(L) + torch.triu(torch.full((n, n), 7.0, dtype=torch.float64, device=dev()), 1)  # junk above
    Bd = conv(B)
    be.trsm_rlt_prefix_(Ld, Bd, act)
    assert_close(Bd, U[rows], 1e-12, 'rows of U')
    Lt = be.transpose(Ld)
    be.trsm_rln_prefix_(Ld, Lt, Bd, act)
    got = Bd.cpu().numpy()
    want = Kinv[rows]
    keep = np.arange(n)[None, :] >= (rows // bs * bs)[:, None]
    err = np.abs(got - want)[keep].max() / np.abs(want).max()
    assert err < 1e-11, err
    # full-row mode
    Bd2 = conv(U[rows])
    be.trsm_rln_prefix_(Ld, Lt, Bd2, np.zeros(len(rows), dtype=np.int64))
    assert_close(Bd2, Kinv[rows], 1e-11, 'full rows of K^-1')


def test_weight_rows_kernel():
    rng = np.random.default_rng(0)
    n, bs, R = 700, 256, 2
    rows = np.concatenate([np.arange(256, 512), np.arange(512, 700)])
    W = rng.standard_normal((len(rows), n))
    beta = rng.standard_normal((R, n))
    c0 = rows // bs * bs
    cols = np.arange(n)[None, :]
    ref = 0.5 * (R * W - beta[:, rows].T @ beta)
    ref = np.where(cols >= c0[:, None] + bs, 2 * ref, ref)
    ref = np.where(cols >= c0[:, None], ref, 0.0)
    Wd = conv(W)
    _be().weight_rows_(Wd, torch.tensor(rows, device=dev()), conv(beta), bs)
    assert_close(Wd, ref, 1e-14, 'weight rows')


@pytest.mark.parametrize('n,r,block', [(1000, 1, 256), (1337, 2, 128), (2048, 1, 512)])
def test_distributed_path_world1_equals_fused_and_oracle(n, r, block):
    import gpflowSlim as gpf
    from oracle import cases
    from oracle import ref_torch as R
    d = 4
    X, Y = cases.synth_gpr(n, d)
    rng = np.random.default_rng(3)
    Y = np.concatenate([Y] + [rng.standard_normal((n, 1)) for _ in range(r - 1)], 1)
    kern = gpf.kernels.RBF(d, ARD=True, lengthscales=2.0)
    m = gpf.models.GPR(conv(X), conv(Y), kern=kern)
    params = [p.unconstrained_tensor for p in m.parameters]
    obj = m.objective
    g = torch.autograd.grad(obj, params)
    from gpflowSlim._backend.dist_gpr import CudaBackend
    gpf.parallel.init(block=block)
    CudaBackend.poison = True       # NaN-fill every uninitialised buffer of the distributed path
    try:
        obj2 = m.objective
        g2 = torch.autograd.grad(obj2, params)
    finally:
        CudaBackend.poison = False
        gpf.parallel.shutdown()
    raw = [torch.tensor(R.softplus_inv(v), dtype=torch.float64, requires_grad=True)
           for v in (1.0, 2.0 * np.ones(d), 0.1)]
    spec = dict(type='rbf', variance=R.softplus_fwd(raw[0]), lengthscales=R.softplus_fwd(raw[1]))
    o = R.gpr_nlml(spec, torch.tensor(X), torch.tensor(Y), R.softplus_fwd(raw[2]))
    go = torch.autograd.grad(o, raw)
    assert_close(obj2, o.detach().numpy(), 1e-8, 'dist objective vs oracle')
    for a, b, c in zip(g2, go, g):
        assert_close(a, b.numpy(), 1e-8, 'dist grad vs oracle')
        assert_close(a, c.cpu().numpy(), 1e-9, 'dist grad vs fused')


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs >= 2 GPUs')
def test_two_rank_nccl_run_matches_single_gpu():
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                          '--master-addr', '127.0.0.1', '--master-port', '29533',
                          os.path.join(ROOT, 'tools', 'dist_check.py'), '--size', '3000'],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert 'DIST_CHECK_OK' in out.stdout
