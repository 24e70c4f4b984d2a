"""world_size > 1 coverage on CPU (gloo): the host logic of the block-row distributed GPR path
(gpflowSlim/_backend/dist_gpr.py) with a torch-CPU stand-in for the CUDA kernels, against the
oracle's NLML + autograd gradient."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import cases
from oracle import ref_torch as R

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _spec(theta, d):
    return dict(type='rbf', variance=theta[0], lengthscales=theta[1:1 + d])


def _oracle(n, d, r, noise, ls):
    X, Y = cases.synth_gpr(n, d)
    rng = np.random.default_rng(5)
    Y = np.concatenate([Y] + [rng.standard_normal((n, 1)) for _ in range(r - 1)], 1)
    th = torch.tensor([1.3] + [ls] * d, dtype=torch.float64, requires_grad=True)
    nz = torch.tensor(noise, dtype=torch.float64, requires_grad=True)
    Yt = torch.tensor(Y, requires_grad=True)
    obj = R.gpr_nlml(_spec(th, d), torch.tensor(X), Yt, nz)
    g = torch.autograd.grad(obj, [th, nz, Yt])
    return X, Y, th.detach(), obj.detach(), [t.detach() for t in g]


def _worker(rank, world, port, n, d, r, block, lookahead, out_q):
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    from dist_cpu_backend import CpuBackend
    from gpflowSlim._backend import dist_gpr
    torch.set_num_threads(2)
    if world > 1:
        dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    X, Y, th, obj, g = _oracle(n, d, r, 0.1, 1.7)

    class Prog(object):
        n_theta = 1 + d
    be = CpuBackend(lambda t: _spec(t, d))
    nlml, dth, dnz, dY = dist_gpr.nlml_and_grad(Prog(), th, 0.1, torch.tensor(X), torch.tensor(Y),
                                                block=block, backend=be, lookahead=lookahead)

    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max())
    errs = [rel(nlml, obj), rel(dth, g[0]), rel(dnz, g[1]), rel(dY, g[2])]
    out_q.put((rank, errs))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n,r,block,lookahead', [(1, 700, 1, 256, True), (2, 900, 2, 128, True),
                                                       (3, 1100, 1, 256, True), (2, 600, 1, 128, False),
                                                       (2, 515, 3, 256, True), (2, 900, 2, 128, 'v2'),
                                                       (4, 1300, 1, 128, 'v2')])
def test_distributed_gpr_host_logic_matches_oracle(world, n, r, block, lookahead):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(i, world, port, n, 3, r, block, lookahead, q)) for i in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errs in res:
        assert max(errs) < 1e-9, (rank, errs)


def _predict_worker(rank, world, port, n, ns, r, block, out_q):
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    from dist_cpu_backend import CpuBackend
    from gpflowSlim._backend import dist_gpr
    torch.set_num_threads(2)
    if world > 1:
        dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    d = 3
    X, Y, th, _, _ = _oracle(n, d, r, 0.1, 1.7)
    Xs = np.random.default_rng(9).standard_normal((ns, d))
    spec = _spec(th, d)
    mo, vo = R.gpr_predict(spec, torch.tensor(X), torch.tensor(Y), torch.tensor(0.1, dtype=torch.float64),
                           torch.tensor(Xs))

    class Prog(object):
        n_theta = 1 + d
    be = CpuBackend(lambda t: _spec(t, d))
    mean, var = dist_gpr.predict(Prog(), th, 0.1, torch.tensor(X), torch.tensor(Y), torch.tensor(Xs),
                                 R.Kdiag(spec, torch.tensor(Xs)), block=block, backend=be)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    out_q.put((rank, [rel(mean, mo), rel(var, vo.reshape(ns, -1)[:, 0])]))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n,ns,r,block', [(2, 700, 53, 2, 128), (3, 640, 5, 1, 256), (4, 600, 2, 1, 128)])
def test_distributed_predict_matches_oracle(world, n, ns, r, block):
    """dist_gpr.predict: the ranks factor together, the test points are dealt out (ragged, also
    fewer test points than ranks), one all-gather joins mean / variance."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_predict_worker, args=(i, world, port, n, ns, r, block, q)) for i in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errs in res:
        assert max(errs) < 1e-9, (rank, errs)


def test_layout_is_a_partition_and_balanced():
    from gpflowSlim._backend.dist_gpr import BlockRowLayout
    lay = BlockRowLayout(32768, 512, 8)
    owned = sorted(b for q in range(8) for b in lay.blocks_of(q))
    assert owned == list(range(lay.nblk))
    # snake order: triangular work (sum of block indices) identical across ranks
    w = [sum(lay.blocks_of(q)) for q in range(8)]
    assert max(w) - min(w) == 0
    inv = lay.inverse_assignment()
    assert sorted(b for m in inv for b in m) == list(range(lay.nblk))
    cost = [sum((lay.n - lay.rows(b)[0]) ** 2 for b in m) for m in inv]
    assert max(cost) / (sum(cost) / 8) < 1.05
    lo, m = lay.rows_below(3, 10)
    offs, nloc = lay.local_offsets(3)
    assert lo == offs[12] and m == nloc - lo      # rank 3 owns 3, 12, 19, 28, ...
