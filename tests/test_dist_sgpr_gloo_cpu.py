"""world_size > 1 coverage on CPU (gloo) of the data-parallel SGPR objective
(gpflowSlim/parallel.py:sgpr_objective_and_grads; SURVEY.md section 8e: "shard N columns of Kuf,
allreduce M x M + M x R"; reference models/sgpr.py:121-156).  Kernels = the torch-CPU test double;
tested is the host logic: ragged shards, the reduced statistics, the two-stage gradient (direct
part replicated, pulled-back part all-reduced).  Every rank must reproduce the single-process
objective and gradients (incl. the inducing inputs Z) on the concatenated data."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Patch(object):
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def _worker(rank, world, port, r, out_q):
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    import cpu_ops_double
    import gpflowSlim as gpf
    from oracle import cases
    torch.set_num_threads(2)
    cpu_ops_double.install(_Patch())
    gpf.settings.device = 'cpu'
    if world > 1:
        dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    n, d, m = 407, 3, 19
    X, Y, Z = cases.synth_svgp(n, d, m, seed=6)
    rng = np.random.default_rng(8)
    Y = np.concatenate([Y] + [rng.standard_normal((n, 1)) for _ in range(r - 1)], 1)
    conv = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))

    def make(Xp, Yp):
        kern = gpf.kernels.RBF(d, ARD=True, lengthscales=1.4) + gpf.kernels.Linear(d, variance=0.3)
        return gpf.models.SGPR(conv(Xp), conv(Yp), kern, Z=Z.copy(), obs_var=0.15,
                               mean_function=gpf.mean_functions.Constant(np.full(r, 0.2)))
    whole = make(X, Y)
    pw = whole.trainable_tensors
    obj = whole.objective
    want = torch.autograd.grad(obj, pw, allow_unused=True)
    bounds = np.linspace(0, n, world + 1).round().astype(int)
    sl = slice(bounds[rank], bounds[rank + 1])
    mine = make(X[sl], Y[sl])
    if world > 1:
        gpf.parallel.init(group=None, backend='gloo')
    got_obj, got = gpf.parallel.sgpr_objective_and_grads(mine)

    def rel(a, b):
        b = torch.zeros_like(a) if b is None else b
        return float((a - b).abs().max() / max(float(b.abs().max()), 1e-300))
    assert len(got) == len(want)
    out_q.put((rank, [rel(got_obj, obj.detach())] + [rel(g, w) for g, w in zip(got, want)]))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize('world,r', [(1, 1), (2, 2), (3, 1)])
def test_sharded_sgpr_equals_single_process(world, r):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(i, world, port, r, q)) for i in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errs in res:
        assert max(errs) < 1e-9, (rank, errs)
