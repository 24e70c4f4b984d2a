"""Times the distributed GPR path (any world size, also 1) phase by phase with CUDA events.
Launch: python tools/dist_time.py --n 16384   or under torchrun."""
import argparse
import math
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'gpflow-slim_b200'))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=16384, dest='n')
    ap.add_argument('--d', type=int, default=8)
    ap.add_argument('--block', type=int, default=512)
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--fused', type=int, default=1)
    ap.add_argument('--lookahead', type=int, default=1)
    args = ap.parse_args()
    import gpflowSlim as gpf
    from bench import synth_gpr
    local = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    gpf.settings.device = dev
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    rank = dist.get_rank() if world > 1 else 0
    n, d = args.n, args.d
    X, Y = synth_gpr(n, d)
    conv = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    kern = gpf.kernels.RBF(d, ARD=True, lengthscales=math.sqrt(d))
    m = gpf.models.GPR(conv(X), conv(Y), kern=kern)
    params = [p.unconstrained_tensor for p in m.parameters]

    def step():
        obj = m.objective
        return obj, torch.autograd.grad(obj, params)

    def timeit(label):
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.reps):
            o, g = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        wall = (time.perf_counter() - t0) / args.reps * 1e3
        if rank == 0:
            print('%s: N=%d world=%d block=%d  %.1f ms/eval (wall %.1f)  %.2f TFLOP/s (N^3 model) obj=%.6f'
                  % (label, n, world, args.block, ms, wall, n ** 3 / ms / 1e9, float(o)), flush=True)
        return o, g

    ref = None
    if args.fused and rank == 0 or (args.fused and world > 1):
        ref = timeit('fused 1-GPU')
    gpf.parallel.init(block=args.block, lookahead=bool(args.lookahead))
    o, g = timeit('distributed')
    from gpflowSlim._backend import dist_gpr
    dist_gpr.TIMER = dist_gpr.PhaseTimer()
    step()
    rep = dist_gpr.TIMER.report()
    dist_gpr.TIMER = None
    if rank == 0:
        print('phases (ms): ' + ', '.join('%s %.1f' % kv for kv in rep.items()), flush=True)
    gpf.parallel.shutdown()
    if ref is not None and rank == 0:
        err = max([abs(float(o) - float(ref[0])) / abs(float(ref[0]))] +
                  [float((a - b).abs().max() / b.abs().max()) for a, b in zip(g, ref[1])])
        print('max rel err dist vs fused: %.2e' % err, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
