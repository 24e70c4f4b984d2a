"""world_size > 1 coverage on CPU (gloo) of the data-parallel SVGP step (SURVEY.md section 8e:
minibatch rows sharded across ranks, Kuu / chol(Kuu) / KL replicated, ONE all-reduce of a flat
[objective, gradients] buffer -- gpflowSlim/parallel.py:svgp_objective_and_grads).  The kernels
are the torch-CPU test double (tests/cpu_ops_double.py); what is tested is the host logic: ragged
shards, the num_data / B scale taken over the GLOBAL batch, the KL / world split, the buffer
packing.  Every rank must reproduce the single-process objective and gradients (including the
inducing inputs Z) of the concatenated batch."""
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


def _worker(rank, world, port, whiten, q_diag, multiclass, out_q):
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
    n, d, m, batch = 900, 4, 24, 203                     # 203 rows: ragged over 2 and 3 ranks
    X, Y, Z = cases.synth_svgp(n, d, m, seed=3)
    rng = np.random.default_rng(4)
    if multiclass:
        Y = rng.integers(0, 3, (n, 1)).astype(np.float64)
        lik, latents = gpf.likelihoods.MultiClass(3), 3
    else:
        lik, latents = gpf.likelihoods.Gaussian(var=0.2), 1
    conv = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    kern = gpf.kernels.RBF(d, ARD=True, lengthscales=1.5)
    model = gpf.models.SVGP(conv(X[:batch]), conv(Y[:batch]), kern, lik, Z=Z.copy(), whiten=whiten,
                            q_diag=q_diag, num_latent=latents, num_data=n)
    with torch.no_grad():
        model._q_mu.unconstrained_tensor.add_(conv(0.3 * rng.standard_normal(tuple(model._q_mu.shape))))
        model._q_sqrt.unconstrained_tensor.add_(conv(0.05 * rng.standard_normal(tuple(model._q_sqrt.shape))))
    params = model.trainable_tensors
    # single-process reference on the whole batch
    obj = model.objective
    want = torch.autograd.grad(obj, params, allow_unused=True)
    # this rank's shard: contiguous, sizes differ by at most one row
    bounds = np.linspace(0, batch, world + 1).round().astype(int)
    sl = slice(bounds[rank], bounds[rank + 1])
    if world > 1:
        gpf.parallel.init(group=None, backend='gloo')
    got_obj, got = gpf.parallel.svgp_objective_and_grads(model, conv(X[:batch][sl]), conv(Y[:batch][sl]), params)

    def rel(a, b):
        b = torch.zeros_like(a) if b is None else b
        return float((a - b).abs().max() / max(float(b.abs().max()), 1e-300))
    errs = [rel(got_obj, obj.detach())] + [rel(g, w) for g, w in zip(got, want)]
    out_q.put((rank, errs))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize('world,whiten,q_diag,multiclass', [(2, True, False, False), (3, False, False, False),
                                                            (2, True, True, True)])
def test_sharded_svgp_step_equals_single_process(world, whiten, q_diag, multiclass):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(i, world, port, whiten, q_diag, multiclass, q)) for i in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errs in res:
        # the multi-class kernel-variance gradient is ~1e-6 of the others (cancellation): held to
        # the north-star tolerance
        assert max(errs) < 1e-8, (rank, errs)
