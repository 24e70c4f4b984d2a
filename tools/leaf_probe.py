"""Runs a few small factorisations (for ncu captures of the leaf / strip kernels) and prints their
CUDA-event times:  python tools/leaf_probe.py [n ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'gpflow-slim_b200'))
sys.path.insert(0, ROOT)


def main():
    from gpflowSlim._backend import lib as L
    from gpflowSlim._backend import ops
    dev = torch.device('cuda', 0)
    sizes = [int(a) for a in sys.argv[1:]] or [128, 512, 1024, 2048]
    h = L.handle_for(dev)
    for n in sizes:
        rng = np.random.default_rng(n)
        A = rng.standard_normal((n, n + 8))
        K = torch.as_tensor(A @ A.T / n + np.eye(n), dtype=torch.float64, device=dev)
        W = torch.empty_like(K)
        U = torch.empty_like(K)

        def potrf():
            W.copy_(K)
            v = L.view(W)
            h.check(h.lib.gps_potrf(h.ptr, v.ref, 0, None))

        def inv():
            vl, vu = L.view(W), L.view(U)
            h.check(h.lib.gps_tri_inv_t(h.ptr, vl.ref, vu.ref))
        res = {}
        for name, fn in (('copy', lambda: W.copy_(K)), ('potrf+copy', potrf), ('tri_inv_t', inv)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / 20 * 1e3
        Lt = torch.linalg.cholesky(K)
        err = float((torch.tril(W) - Lt).abs().max() / Lt.abs().max())
        print('n=%d: potrf %.1f us, tri_inv_t %.1f us (copy %.1f us), max rel err vs torch %.1e'
              % (n, res['potrf+copy'] - res['copy'], res['tri_inv_t'], res['copy'], err), flush=True)


if __name__ == '__main__':
    main()
