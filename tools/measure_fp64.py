"""Measures, on the GPU box, the FP64 denominators MEASURED_PEAKS.json lacks -- cuBLAS DGEMM
(burst / sustained, same method as the driver's bf16 figure), and the vendor kernels to beat:
cuSOLVER potrf, cuBLAS trsm / syrk via torch -- and times this library's GEMM / POTRF beside
them.  Writes gpurun_out/fp64_peak.json (copy the summary to profiles/)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'gpflow-slim_b200'))
sys.path.insert(0, ROOT)


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best, tot = 1e30, 0.0
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best = min(best, ms)
        tot += ms
    return best, tot / reps


def main():
    from gpflowSlim._backend import ops
    dev = torch.device('cuda', 0)
    out = {'gpu': torch.cuda.get_device_name(0), 'torch': torch.__version__}
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty_like(a)
    best, avg = timeit(lambda: torch.matmul(a, b.t(), out=c), reps=10)
    out['cublas_dgemm_8192_tflops_burst'] = 2.0 * n ** 3 / best / 1e9
    t0 = time.time()
    cnt = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(5):
            torch.matmul(a, b.t(), out=c)
        cnt += 5
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    out['cublas_dgemm_8192_tflops_sustained'] = 2.0 * n ** 3 * cnt / e0.elapsed_time(e1) / 1e9
    best, avg = timeit(lambda: ops.gemm_nt(a, b, out=c, beta=0.0), reps=10)
    out['ours_dgemm_nt_8192_tflops_burst'] = 2.0 * n ** 3 / best / 1e9
    for nn in (2048, 4096, 8192, 16384):
        g = torch.randn(nn, nn + 16, dtype=torch.float64, device=dev)
        s = g @ g.t() / nn + 0.5 * torch.eye(nn, dtype=torch.float64, device=dev)
        del g
        best, _ = timeit(lambda: torch.linalg.cholesky(s), reps=3, warm=1)
        out['cusolver_potrf_%d_tflops' % nn] = nn ** 3 / 3.0 / best / 1e9
        out['cusolver_potrf_%d_ms' % nn] = best
        best, _ = timeit(lambda: ops.potrf(s, zero_upper=False, check=False), reps=3, warm=1)
        out['ours_potrf_%d_tflops' % nn] = nn ** 3 / 3.0 / best / 1e9
        out['ours_potrf_%d_ms' % nn] = best
        L = torch.linalg.cholesky(s)
        rhs = torch.randn(1024, nn, dtype=torch.float64, device=dev)
        best, _ = timeit(lambda: torch.linalg.solve_triangular(L, rhs.t(), upper=False), reps=3, warm=1)
        out['cublas_trsm_%dx1024_ms' % nn] = best
        best, _ = timeit(lambda: ops.trsm_rlt_(L, rhs.clone()), reps=3, warm=1)
        out['ours_trsm_%dx1024_ms' % nn] = best
        best, _ = timeit(lambda: torch.cholesky_inverse(L), reps=3, warm=1)
        out['cusolver_potri_%d_ms' % nn] = best
        del s, L, rhs
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'fp64_peak.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
