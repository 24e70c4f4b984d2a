"""Host-side cost of every operation the distributed factorisation issues (under torchrun): wraps
the backend and communicator methods with perf_counter timers and prints, per rank, the calls on
which the HOST spent the most time (a call that blocks until the GPU catches up shows up here)."""
import argparse
import collections
import math
import os
import sys
import time

os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'gpflow-slim_b200'))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=32768, dest='n')
    ap.add_argument('--block', type=int, default=512)
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
    gpf.parallel.init(block=args.block)

    def step():
        obj = m.objective
        return obj, torch.autograd.grad(obj, params)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    acc = collections.defaultdict(lambda: [0, 0.0, 0.0])

    def wrap(cls, name):
        fn = getattr(cls, name)

        def timed(*a, **k):
            t0 = time.perf_counter()
            try:
                return fn(*a, **k)
            finally:
                dt = time.perf_counter() - t0
                e = acc[cls.__name__ + '.' + name]
                e[0] += 1
                e[1] += dt
                e[2] = max(e[2], dt)
        setattr(cls, name, timed)
    for name in ('potrf_', 'trsm_rlt_', 'syrk_lower_', 'gemm_rowmap_', 'copy_', 'zero_', 'unpack_rows_', 'transpose_into',
                 'record', 'wait', 'gram_rows', 'trsm_rlt_prefix_', 'trsm_rln_prefix_', 'weight_rows_', 'gram_bwd'):
        wrap(dist_gpr.CudaBackend, name)
    for name in ('broadcast', 'all_gather', 'all_reduce_sum'):
        wrap(dist_gpr._Comm, name)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    step()
    host_ms = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    for r in range(world):
        if world > 1:
            dist.barrier()
        if r == rank and r < 3:
            print('rank %d: host time of one step %.1f ms' % (rank, host_ms))
            for k, (n, tot, mx) in sorted(acc.items(), key=lambda kv: -kv[1][1])[:12]:
                print('   %-32s calls %5d  total %8.2f ms  mean %7.1f us  max %8.1f us' % (k, n, tot * 1e3, tot / n * 1e6, mx * 1e6))
            sys.stdout.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
