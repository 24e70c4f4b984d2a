"""Secondary measurements of SURVEY.md section 8(d) on one B200 (or, for SVGP, under torchrun on
several): one JSON line per workload, CUDA-event timed after warm-up.

    python tools/bench_secondary.py --what c2      # GPR ARD-RBF N=8192 D=8: NLML+grad, predict_f
    python tools/bench_secondary.py --what c3      # NKN GPR N=16384 D=8 (6 primitives, 5 layers)
    python tools/bench_secondary.py --what c4      # SVGP N=1M D=16 M=1024 B=8192: ELBO+grad+Adam
    python tools/bench_secondary.py --what potrf   # Cholesky TFLOP/s, N = 4096 .. 32768

Not the headline bench (bench.py); the lines are kept under profiles/.
"""
import argparse
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'gpflow-slim_b200'))
sys.path.insert(0, ROOT)
from bench import fp64_peak, synth_gpr  # noqa: E402


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def gpr_lines(gpf, dev, kern, n, d, tag, reps):
    X, Y = synth_gpr(n, d)
    conv = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    m = gpf.models.GPR(conv(X), conv(Y), kern=kern)
    params = [p.unconstrained_tensor for p in m.parameters]
    Xs = conv(np.random.default_rng(1).standard_normal((1024, d)))

    def step():
        obj = m.objective
        return obj, torch.autograd.grad(obj, params)

    def pred():
        with torch.no_grad():
            return m.predict_f(Xs)
    ms, (obj, _) = timed(step, reps)
    pms, _ = timed(pred, reps)
    pk = fp64_peak()['burst']
    return [{'workload': tag, 'metric': 'GPR NLML+grad evals/s', 'value': 1e3 / ms, 'ms': ms,
             'tflops_n3_model': n ** 3 / ms / 1e9, 'frac_of_fp64_peak': n ** 3 / ms / 1e9 / pk,
             'objective': float(obj), 'n_params': int(sum(p.numel() for p in params))},
            {'workload': tag, 'metric': 'predict_f latency (1024 test points)', 'ms': pms,
             'tflops': (n ** 3 / 3 + 2.0 * n * n * 1024) / pms / 1e9}]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--what', default='c2')
    ap.add_argument('--reps', type=int, default=5)
    args = ap.parse_args()
    import gpflowSlim as gpf
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    rank = int(os.environ.get('RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    gpf.settings.device = dev
    if os.environ.get('GPSLIM_GRAM_IMPL'):      # e.g. 2 = experimental smem-accumulator Gram backward
        from gpflowSlim._backend import lib as _L
        _L.handle_for(dev).set_option('gram_impl', int(os.environ['GPSLIM_GRAM_IMPL']))
    lines = []
    if args.what == 'c2':
        kern = gpf.kernels.RBF(8, ARD=True, lengthscales=math.sqrt(8))
        lines = gpr_lines(gpf, dev, kern, 8192, 8, 'C2 GPR ARD-RBF N=8192 D=8 fp64', args.reps)
    elif args.what == 'c3':
        from oracle import cases      # the C3 topology is defined once, next to the parity cases
        kern = cases.nkn_c3_kernel(gpf, 8)
        lines = gpr_lines(gpf, dev, kern, 16384, 8, 'C3 NKN GPR N=16384 D=8 fp64 (6 primitives, '
                          'Linear6-8/Product/Linear4-4/Product/Linear2-1)', args.reps)
    elif args.what == 'potrf':
        import ctypes  # noqa: F401
        from gpflowSlim._backend import lib as L
        h = L.handle_for(dev)
        pk = fp64_peak()['burst']
        for n in (4096, 8192, 16384, 32768):
            X, _ = synth_gpr(n, 8)
            with torch.no_grad():
                K = gpf.kernels.RBF(8, ARD=True, lengthscales=math.sqrt(8)).K(
                    torch.as_tensor(X, device=dev))
                K.diagonal().add_(0.1)
            A = torch.empty_like(K)

            def f():
                A.copy_(K)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                va = L.view(A)
                e0.record()
                h.check(h.lib.gps_potrf(h.ptr, va.ref, 0, None))
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1)
            f()
            ms = float(np.mean([f() for _ in range(3)]))
            tf = n ** 3 / 3.0 / ms / 1e9
            lines.append({'workload': 'POTRF N=%d fp64 (K + 0.1 I, ARD-RBF D=8)' % n, 'ms': ms,
                          'tflops': tf, 'frac_of_fp64_peak': tf / pk})
            del K, A
    elif args.what == 'c4':
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group('nccl', device_id=dev)
            gpf.parallel.init()
        n, d, M, B = 1000000, 16, 1024, 8192
        rng = np.random.default_rng(0)
        X = rng.standard_normal((n, d))
        Y = np.sin(X.sum(1, keepdims=True) / 4.0) + 0.1 * rng.standard_normal((n, 1))
        Z = X[np.random.default_rng(2).permutation(n)[:M]].copy()
        Xd, Yd = torch.as_tensor(X, device=dev), torch.as_tensor(Y, device=dev)
        kern = gpf.kernels.RBF(d, ARD=True, lengthscales=4.0)
        m = gpf.models.SVGP(Xd[:B], Yd[:B], kern, gpf.likelihoods.Gaussian(var=0.1), Z=Z, num_data=n)
        params = m.trainable_tensors
        opt = gpf.training.AdamOptimizer(1e-3)
        state = {'i': 0}
        bl = B // world

        def step():
            i0 = (state['i'] * B) % (n - B)
            state['i'] += 1
            if world > 1:
                sl = slice(i0 + rank * bl, i0 + (rank + 1) * bl)
                obj, grads = gpf.parallel.svgp_objective_and_grads(m, Xd[sl], Yd[sl], params)
            else:
                m.X, m.Y = Xd[i0:i0 + B], Yd[i0:i0 + B]
                obj = m.objective
                grads = torch.autograd.grad(obj, params)
            opt.apply_gradients(zip(grads, params))
            return obj
        ms, obj = timed(step, max(args.reps, 20), warm=5)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        flop = 5.2e10     # forward + backward model, K = 1
        lines = [{'workload': 'C4 SVGP Gaussian N=1M D=16 M=1024 B=8192 whiten, full q_sqrt, fp64: '
                              'ELBO + grad (Z, q_mu, q_sqrt, theta) + Adam step', 'n_gpus': world,
                  'metric': 'SVGP steps/s', 'value': 1e3 / ms, 'ms': ms, 'tflops_model': flop / ms / 1e9,
                  'objective_last': float(obj)}]
        if world == 1:
            # the same step captured once in a CUDA graph and replayed (SURVEY.md 8d timing method)
            m2 = gpf.models.SVGP(Xd[:B], Yd[:B], gpf.kernels.RBF(d, ARD=True, lengthscales=4.0),
                                 gpf.likelihoods.Gaussian(var=0.1), Z=Z, num_data=n)
            gstep = gpf.training.GraphedStep(m2, Xd[:B], Yd[:B], learning_rate=1e-3)
            st2 = {'i': 0}

            def graphed():
                i0 = (st2['i'] * B) % (n - B)
                st2['i'] += 1
                return gstep(Xd[i0:i0 + B], Yd[i0:i0 + B])
            gms, gobj = timed(graphed, max(args.reps, 20), warm=5)
            lines.append(dict(lines[0], metric='SVGP steps/s (CUDA-graph replay)', value=1e3 / gms, ms=gms,
                              tflops_model=flop / gms / 1e9, objective_last=float(gobj)))
        if world > 1:
            dist.destroy_process_group()
    if rank == 0:
        for l in lines:
            print(json.dumps(l))


if __name__ == '__main__':
    main()
