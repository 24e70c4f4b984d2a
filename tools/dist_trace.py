"""Per-panel timeline of the distributed factorisation's look-ahead schedule (under torchrun):
for every panel the GPU time (ms since the streams fork) at which each stage finished, and the
host time at which the panel was issued.  Tells chain-bound from bulk-bound from host-bound.

    torchrun ... tools/dist_trace.py --size 32768 [--block 512]
"""
import argparse
import json
import math
import os
import sys

os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', os.environ.get('GPSLIM_MAXCONN', '32'))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'gpflow-slim_b200'))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=32768, dest='n')
    ap.add_argument('--block', type=int, default=512)
    ap.add_argument('--out', default='gpurun_out/dist_trace')
    args = ap.parse_args()
    import gpflowSlim as gpf
    from bench import synth_gpr
    from gpflowSlim._backend import dist_gpr
    local = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    gpf.settings.device = dev
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    rank = dist.get_rank() if world > 1 else 0
    X, Y = synth_gpr(args.n, 8)
    conv = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    m = gpf.models.GPR(conv(X), conv(Y), kern=gpf.kernels.RBF(8, ARD=True, lengthscales=math.sqrt(8)))
    params = [p.unconstrained_tensor for p in m.parameters]
    gpf.parallel.init(block=args.block, lookahead=os.environ.get('GPSLIM_SCHEDULE', 'v2'))

    def step():
        obj = m.objective
        return obj, torch.autograd.grad(obj, params)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dist_gpr.TRACE = {}
    step()
    tr = dict(dist_gpr.TRACE)
    dist_gpr.TRACE = None
    json.dump(tr, open('%s_world%d_rank%d.json' % (args.out, world, rank), 'w'))
    if rank == 0:
        nblk = len(tr['L'])
        print('factor total %.1f ms; host issued the last panel at %.1f ms' % (tr['total_ms'], tr['host_issue_ms'][nblk - 1]))
        print('%4s %8s | %8s %8s %8s %8s %8s %8s %8s' % ('k', 'host', 'L', 'top', 'solve', 'gathered', 'narrow', 'col', 'rest'))
        for k in range(nblk):
            g = lambda f: ('%8.2f' % tr[f][k]) if k in tr[f] else '       -'
            print('%4d %8.2f | %s %s %s %s %s %s %s' % (k, tr['host_issue_ms'][k], g('L'), g('top'), g('solve'), g('gathered'),
                                                       g('narrow'), g('col'), g('rest')))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
