"""Quick GPU check of the GEMM kernels: TMA (impl 0) and cp.async (impl 2) against torch fp64
on a few shapes, then 8192^3 timings beside cuBLAS DGEMM."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'gpflow-slim_b200'))
from gpflowSlim._backend import ops  # noqa: E402
from gpflowSlim._backend.lib import handle_for  # noqa: E402

dev = torch.device('cuda', 0)
h = handle_for(dev)
g = torch.Generator(device='cuda').manual_seed(0)
for impl in (0, 2):
    h.set_option('gemm_impl', impl)
    worst = 0.0
    for (m, n, k) in [(128, 128, 16), (128, 128, 128), (257, 513, 384), (200, 300, 50), (385, 129, 4098),
                      (1000, 2, 1000), (2048, 2048, 2048), (3000, 2500, 144), (4100, 1700, 16)]:
        A = torch.randn(m, k, dtype=torch.float64, device=dev, generator=g)
        B = torch.randn(n, k, dtype=torch.float64, device=dev, generator=g)
        C = torch.randn(m, n, dtype=torch.float64, device=dev, generator=g)
        ref = 0.7 * A @ B.t() - 0.3 * C
        out = ops.gemm_nt(A, B, alpha=0.7, beta=-0.3, out=C.clone())
        torch.cuda.synchronize()
        e = float((out - ref).abs().max() / ref.abs().max())
        worst = max(worst, e)
        print('impl %d  %5d x %5d x %5d  rel err %.2e' % (impl, m, n, k, e), flush=True)
    assert worst < 1e-13, worst
n = 8192
A = torch.randn(n, n, dtype=torch.float64, device=dev)
B = torch.randn(n, n, dtype=torch.float64, device=dev)
C = torch.empty_like(A)


def timeit(f, reps=5):
    f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for impl, name in ((0, 'TMA+mbarrier'), (2, 'cp.async')):
    h.set_option('gemm_impl', impl)
    ms = timeit(lambda: ops.gemm_nt(A, B, out=C, beta=0.0))
    print('ours %-14s 8192^3: %.2f ms  %.2f TFLOP/s' % (name, ms, 2 * n ** 3 / ms / 1e9), flush=True)
    for k in (128, 256, 512, 1024, 2048):
        Ak, Bk = A[:, :k], B[:, :k]
        ms = timeit(lambda: ops.gemm_nt(Ak, Bk, out=C, beta=1.0, alpha=-1.0))
        print('ours %-14s 8192x8192x%d (beta=1): %.3f ms  %.2f TFLOP/s' % (name, k, ms, 2 * n * n * k / ms / 1e9), flush=True)
h.set_option('gemm_impl', 0)
ms = timeit(lambda: torch.matmul(A, B.t(), out=C))
print('cuBLAS DGEMM       8192^3: %.2f ms  %.2f TFLOP/s' % (ms, 2 * n ** 3 / ms / 1e9))
