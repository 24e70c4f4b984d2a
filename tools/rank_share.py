"""What ONE rank of a P-rank distributed GPR evaluation computes, timed on a single GPU (no
communication): the pieces of `_backend/dist_gpr.py` that need no exchange are run with the
layout of rank r of P.  Separates "kernel efficiency at the per-rank shapes" from "latency of the
serial chain + collectives", which only a P-GPU run shows.

    python tools/rank_share.py --size 32768 --world 8 --rank 0

Prints one JSON line: chain pieces (diagonal-block factorisation, block inverse, top-block solve,
narrow updates) in microseconds, the rank's bulk trailing updates and its inverse-row phases in
milliseconds, each with the TFLOP/s it corresponds to.
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


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=32768, dest='n')
    ap.add_argument('--d', type=int, default=8)
    ap.add_argument('--world', type=int, default=8)
    ap.add_argument('--rank', type=int, default=0)
    ap.add_argument('--block', type=int, default=512)
    ap.add_argument('--what', default='chain,bulk,inverse')
    args = ap.parse_args()
    import gpflowSlim as gpf
    from bench import synth_gpr
    from gpflowSlim._backend import dist_gpr as D
    from gpflowSlim._backend import lib as L
    from gpflowSlim._backend import ops
    dev = torch.device('cuda', 0)
    gpf.settings.device = dev
    n, d, P, r, bs = args.n, args.d, args.world, args.rank, args.block
    be = D.CudaBackend(dev)
    lay = D.BlockRowLayout(n, bs, P)
    out = {'n': n, 'world': P, 'rank': r, 'block': bs}
    what = args.what.split(',')
    conv = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    X, Y = synth_gpr(n, d)
    X, Y = conv(X), conv(Y)
    kern = gpf.kernels.RBF(d, ARD=True, lengthscales=math.sqrt(d))
    prog = kern.program()
    theta = prog.theta(dev).detach()

    if 'chain' in what:
        # one diagonal block of the real matrix, already "updated" (K + noise I is SPD itself)
        Kb = be.empty(bs, bs)
        be.gram_rows(prog, theta, X[:bs], X[:bs], Kb)
        Kb.diagonal().add_(0.1)
        Dm = torch.empty_like(Kb)
        Lkk = ops.potrf(Kb)
        chain = {}

        def f_potrf():
            Dm.copy_(Kb)
            be.potrf_(Dm)
        t_copy = timed(lambda: Dm.copy_(Kb), reps=20)
        chain['potrf_%d_us' % bs] = (timed(f_potrf, reps=20) - t_copy) * 1e3
        Ub = torch.empty_like(Lkk)
        h = L.handle_for(dev)

        def f_inv():
            vl, vu = L.view(Lkk), L.view(Ub)
            h.check(h.lib.gps_tri_inv_t(h.ptr, vl.ref, vu.ref))
        chain['tri_inv_%d_us' % bs] = timed(f_inv, reps=20) * 1e3
        for rows in (bs, n // P):
            Pn = conv(np.random.default_rng(0).standard_normal((rows, bs)))
            chain['trsm_rows%d_us' % rows] = timed(lambda: be.trsm_rlt_(Lkk, Pn), reps=20) * 1e3
            C = torch.empty(rows, bs, dtype=torch.float64, device=dev)
            Tl = Ub.t().contiguous()
            chain['gemm_with_inverse_rows%d_us' % rows] = timed(
                lambda: ops.gemm_nt(Pn, Tl, out=C, b_tri=L.TRI_LOWER), reps=20) * 1e3
        Xt = conv(np.random.default_rng(1).standard_normal((bs, bs)))
        Cs = torch.zeros(bs, bs, dtype=torch.float64, device=dev)
        chain['syrk_%d_us' % bs] = timed(lambda: ops.gemm_nt(Xt, Xt, alpha=-1.0, beta=1.0, out=Cs, c_uplo=1),
                                         reps=20) * 1e3
        out['chain'] = chain

    ld = D._round_up(n, 16)
    if 'bulk' in what:
        offs, nloc = lay.local_offsets(r)
        Aloc = torch.randn(nloc + 1, ld, dtype=torch.float64, device=dev)
        Lf = torch.randn(n, ld, dtype=torch.float64, device=dev)
        grow = torch.full((nloc + 1,), D.NEVER, dtype=torch.int64, device=dev)
        for b in lay.blocks_of(r):
            r0, r1 = lay.rows(b)
            grow[offs[b]:offs[b] + r1 - r0] = torch.arange(r0, r1, dtype=torch.int64, device=dev)
        flops = 0.0
        for k in range(lay.nblk - 1):
            k0, k1 = lay.rows(k)
            for b in lay.blocks_of(r):
                if b > k:
                    g = np.arange(*lay.rows(b))
                    flops += 2.0 * (k1 - k0) * float(np.clip(g - k1 + 1, 0, None).sum())

        def f_bulk(split):
            for k in range(lay.nblk - 1):
                k0, k1 = lay.rows(k)
                lo, mrows = lay.rows_below(r, k)
                if split:
                    for (c_lo, c_hi) in ((k1, min(n, k1 + bs)), (min(n, k1 + bs), min(n, k1 + 2 * bs)),
                                         (min(n, k1 + 2 * bs), n)):
                        if c_hi > c_lo:
                            be.gemm_rowmap_(Aloc[lo:, k0:k1], Lf[c_lo:c_hi, k0:k1], Aloc[lo:, c_lo:c_hi], grow[lo:], c_lo)
                else:
                    be.gemm_rowmap_(Aloc[lo:, k0:k1], Lf[k1:n, k0:k1], Aloc[lo:, k1:n], grow[lo:], k1)
        for split in (0, 1):
            ms = timed(lambda: f_bulk(split), reps=2)
            out['bulk_updates_%s' % ('3_launches_per_panel' if split else '1_launch_per_panel')] = {
                'ms': ms, 'tflops': flops / ms / 1e9}
        del Aloc, Lf

    if 'inverse' in what:
        # the real factor: K + noise I -> L in place (fused library path), L^T by the transpose kernel
        A = be.empty(n, ld)
        bsz = 4096
        for r0 in range(0, n, bsz):
            r1 = min(n, r0 + bsz)
            be.gram_rows(prog, theta, X[r0:r1], X[:r1], A[r0:r1, :r1])
        A[:, :n].diagonal().add_(0.1)
        Lsq = A[:, :n]
        out['potrf_full_ms'] = timed(lambda: be.potrf_(Lsq), reps=1, warm=0)
        Lt = be.empty(n, ld)[:, :n]
        be.transpose_into(Lsq, Lt)
        alpha_t = torch.randn(1, n, dtype=torch.float64, device=dev)
        D.TIMER = D.PhaseTimer()
        acc = {}
        reps = 2
        for it in range(reps + 1):
            D._mark('start')
            D.inverse_rows_and_contract(prog, theta, X, lay, r, Lsq, Lt, alpha_t, be)
            D._mark('contract')
            rep = D.TIMER.report()
            if it:
                for k, v in rep.items():
                    acc[k] = acc.get(k, 0.0) + v / reps
        D.TIMER = None
        share = sum((n - lay.rows(b)[0]) ** 2 * float(lay.rows(b)[1] - lay.rows(b)[0])
                    for b in lay.inverse_assignment()[r])
        out['inverse'] = {k: {'ms': v} for k, v in acc.items()}
        for k in ('rows_of_U', 'rows_of_Kinv'):
            out['inverse'][k]['tflops'] = share / acc[k] / 1e9
    print(json.dumps(out))


if __name__ == '__main__':
    main()
