"""Profiling driver: one warm-up and one profiled fused GPR NLML+grad evaluation (ARD-RBF, or the
NKN network of config C3 with --what nkn; or a bare GEMM / POTRF) so that `ncu --profile-from-start off` captures exactly one step."""
import argparse
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'gpflow-slim_b200'))
sys.path.insert(0, ROOT)
from bench import synth_gpr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--what', default='gpr')
    ap.add_argument('--n', type=int, default=4096)
    ap.add_argument('--d', type=int, default=8)
    args = ap.parse_args()
    import gpflowSlim as gpf
    from gpflowSlim._backend import ops
    dev = torch.device('cuda', 0)
    n, d = args.n, args.d
    if args.what == 'gpr':
        X, Y = synth_gpr(n, d)
        kern = gpf.kernels.RBF(d, ARD=True, lengthscales=math.sqrt(d))
        m = gpf.models.GPR(torch.tensor(X, device=dev), torch.tensor(Y, device=dev), kern=kern)
        params = [p.unconstrained_tensor for p in m.parameters]

        def step():
            obj = m.objective
            torch.autograd.grad(obj, params)
    elif args.what == 'nkn':
        # the BASELINE C3 network through the fused GPR objective: gram_fwd_nkn_kernel / gram_bwd_nkn_kernel
        from bench import nkn_c3_kernel
        X, Y = synth_gpr(n, d)
        m = gpf.models.GPR(torch.tensor(X, device=dev), torch.tensor(Y, device=dev), kern=nkn_c3_kernel(gpf, d))
        params = [p.unconstrained_tensor for p in m.parameters]

        def step():
            obj = m.objective
            torch.autograd.grad(obj, params)
    elif args.what == 'gemm':
        a = torch.randn(n, n, dtype=torch.float64, device=dev)
        b = torch.randn(n, n, dtype=torch.float64, device=dev)
        c = torch.empty_like(a)

        def step():
            ops.gemm_nt(a, b, out=c, beta=0.0)
    else:
        g = torch.randn(n, n + 16, dtype=torch.float64, device=dev)
        s = g @ g.t() / n + 0.5 * torch.eye(n, dtype=torch.float64, device=dev)

        def step():
            ops.potrf(s, zero_upper=False, check=False)
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == '__main__':
    main()
