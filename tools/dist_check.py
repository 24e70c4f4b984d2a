"""Launched under torchrun with >= 2 ranks: distributed GPR objective + gradient (composed
Matern + Linear kernel, and the NKN topology of BASELINE config C3), distributed predict_f, the
sharded SVGP step and the sharded SGPR objective, each against the single-GPU values computed on
the same rank.  The GPR check is repeated (--repeat) with NaN-poisoned buffers: the look-ahead
schedule must give the same bits every time.  Prints DIST_CHECK_OK."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'gpflow-slim_b200'))
sys.path.insert(0, ROOT)


def rel(a, b):
    a, b = a.detach().double().reshape(-1), b.detach().double().reshape(-1)
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-300))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=3000, dest='n')
    ap.add_argument('--block', type=int, default=256)
    ap.add_argument('--repeat', type=int, default=5)
    args = ap.parse_args()
    import gpflowSlim as gpf
    from bench import synth_gpr
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    gpf.settings.device = dev
    dist.init_process_group('nccl', device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    n, d = args.n, 6
    X, Y = synth_gpr(n, d)
    conv = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    kern = gpf.kernels.Matern52(d, ARD=True, lengthscales=2.0) + gpf.kernels.Linear(d, variance=0.2)
    m = gpf.models.GPR(conv(X), conv(Y), kern=kern)
    params = [p.unconstrained_tensor for p in m.parameters]
    obj = m.objective
    g = torch.autograd.grad(obj, params)
    from gpflowSlim._backend.dist_gpr import CudaBackend
    Xnew = conv(np.random.default_rng(1).standard_normal((333, d)))
    with torch.no_grad():
        mu1, var1 = m.predict_f(Xnew)
    gpf.parallel.init(block=args.block)
    CudaBackend.poison = True       # NaN-fill uninitialised buffers: nothing unwritten may be read
    first = None
    for rep in range(args.repeat):
        obj2 = m.objective
        g2 = torch.autograd.grad(obj2, params)
        errs = [rel(obj2, obj)] + [rel(a, b) for a, b in zip(g2, g)]
        # every rank must hold the same answer
        t = torch.stack([obj2.detach()] + [x.reshape(-1)[0] for x in g2])
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        spread = float(((hi - lo).abs() / hi.abs().clamp(min=1e-300)).max())
        assert max(errs) < 1e-8 and spread < 1e-12, (rep, errs, spread)
        if first is None:
            first = t.clone()
        assert torch.equal(first, t), 'run %d differs from run 0: a race in the schedule' % rep
    print('rank %d GPR dist vs single (%d runs, bit-identical): max rel err %.2e, spread over ranks %.2e'
          % (rank, args.repeat, max(errs), spread), flush=True)
    with torch.no_grad():
        mu2, var2 = m.predict_f(Xnew)
    CudaBackend.poison = False
    perr = [rel(mu2, mu1), rel(var2, var1)]
    print('rank %d predict_f dist vs single: max rel err %.2e' % (rank, max(perr)), flush=True)
    assert max(perr) < 1e-8, perr
    gpf.parallel.shutdown()

    # ---- the NKN topology (BASELINE config C3) on the distributed path
    from bench import nkn_c3_kernel
    nn = min(n, 2048)
    mk = gpf.models.GPR(conv(X[:nn, :]), conv(Y[:nn]), kern=nkn_c3_kernel(gpf, d))
    pk = [p.unconstrained_tensor for p in mk.parameters]
    o1 = mk.objective
    g1 = torch.autograd.grad(o1, pk)
    gpf.parallel.init(block=args.block)
    o2 = mk.objective
    g2 = torch.autograd.grad(o2, pk)
    gpf.parallel.shutdown()
    errs = [rel(o2, o1)] + [rel(a, b) for a, b in zip(g2, g1)]
    print('rank %d NKN GPR dist vs single: max rel err %.2e' % (rank, max(errs)), flush=True)
    assert max(errs) < 1e-8, errs

    # ---- SVGP: minibatch sharded over ranks
    from oracle import cases
    nb, M = 512 * world, 64
    Xs, Ys, Z = cases.synth_svgp(nb, 5, M, seed=0)
    sv = gpf.models.SVGP(conv(Xs), conv(Ys), gpf.kernels.RBF(5, ARD=True, lengthscales=2.0),
                         gpf.likelihoods.Gaussian(var=0.1), Z=Z.copy(), num_data=100000)
    ps = sv.trainable_tensors
    o1 = sv.objective
    g1 = torch.autograd.grad(o1, ps)
    sl = slice(rank * 512, (rank + 1) * 512)
    gpf.parallel.init()
    o2, g2 = gpf.parallel.svgp_objective_and_grads(sv, conv(Xs[sl]), conv(Ys[sl]), ps)
    gpf.parallel.shutdown()
    errs = [rel(o2, o1)] + [rel(a, b) for a, b in zip(g2, g1)]
    print('rank %d SVGP sharded vs single: max rel err %.2e' % (rank, max(errs)), flush=True)
    assert max(errs) < 1e-8, errs

    # ---- SGPR: data rows sharded over ranks
    ns = 700 * world + 13
    Xg, Yg, Zg = cases.synth_svgp(ns, 5, 48, seed=1)

    def make(Xp, Yp):
        return gpf.models.SGPR(conv(Xp), conv(Yp), gpf.kernels.RBF(5, ARD=True, lengthscales=2.0), Z=Zg.copy(),
                               obs_var=0.2)
    whole = make(Xg, Yg)
    o1 = whole.objective
    g1 = torch.autograd.grad(o1, whole.trainable_tensors)
    bounds = np.linspace(0, ns, world + 1).round().astype(int)
    mine = make(Xg[bounds[rank]:bounds[rank + 1]], Yg[bounds[rank]:bounds[rank + 1]])
    gpf.parallel.init()
    o2, g2 = gpf.parallel.sgpr_objective_and_grads(mine)
    gpf.parallel.shutdown()
    errs = [rel(o2, o1)] + [rel(a, b) for a, b in zip(g2, g1)]
    print('rank %d SGPR sharded vs single: max rel err %.2e' % (rank, max(errs)), flush=True)
    assert max(errs) < 1e-8, errs
    dist.barrier()
    if rank == 0:
        print('DIST_CHECK_OK', flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
