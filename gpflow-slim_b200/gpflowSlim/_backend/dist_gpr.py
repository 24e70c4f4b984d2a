"""Block-row distributed GPR objective + gradient: one process per GPU, NCCL over NVLink.

What is distributed (reference call sites: models/gpr.py:55-72, densities.py:73-95 and the
TensorFlow gradient of tf.cholesky taken by optimizer.minimize, examples/gpr.py:53-54):

  K + noise I  ->  L  ->  alpha = L^-1 (Y - m)  ->  NLML,
  dNLML/dtheta = sum_ij W_ij dK_ij/dtheta,  W = 1/2 (R K^-1 - beta beta^T),  beta = L^-T alpha.

Layout.  The N x N matrix is cut into block rows of `block` rows (a multiple of 128).  Block
row b belongs to rank `owner(b)` in SNAKE order (0..P-1, P-1..0, ...), which balances the
triangular work of a lower factorisation to within one block.  A rank stores its block rows
contiguously (ascending global order) in one local row-major matrix, followed by the R rows of
(Y - m)^T that ride along every panel solve and leave alpha^T behind (replicated, no TRSV).

Factorisation (right-looking, one exchange per block column k):
  owner(k) factors the diagonal block   -> broadcast  (block^2 doubles)
  every rank solves its rows of the panel against it (DMMA TRSM)
  all-gather of the solved panel        -> every rank now holds column k of L
  every rank updates its own rows of the trailing matrix with ONE masked DMMA GEMM
  (gps_gemm_nt_rowmap: the lower-triangle mask follows the global index of each local row).
Communication per rank: N^2/2 doubles in total, independent of P -- NVSwitch gives every GPU
the same bandwidth to every peer, so the 1-D layout costs no more than a 2-D one at P <= 8
and keeps every GEMM large.  The gathered panels are kept: at the end EVERY rank owns all of L
(N^2 doubles -- 8 GiB at N = 32768, 32 GiB at N = 65536 of the 180 GB).

Gradient.  With L replicated the inverse needs NO further communication: the block rows of
K^-1 are dealt out (largest-first by their (N-c)^2 cost), and each rank computes
  rows of U = L^-T   (gps_trsm_rlt_prefix on rows of the identity),
  rows of K^-1 = U L^-1, columns >= the row block only (gps_trsm_rln_prefix),
at the true flop count (N^3/3 + N^3/3 over all ranks), then contracts its rows of W with
dK/dtheta (Gram tiles recomputed on the fly).  One all-reduce of n_theta + 1 doubles ends
the step.

The arithmetic is behind a small `backend` object: `CudaBackend` (the product: C-ABI calls
into libgpslim_b200.so) -- tests inject a torch-CPU stand-in to exercise this host logic
under gloo without a GPU.
"""
import ctypes
import math

import torch

F64 = torch.float64
NEVER = 1 << 60           # row limit of the ride-along rows: never masked


def _round_up(x, m):
    return (x + m - 1) // m * m


class PhaseTimer(object):
    """Optional CUDA-event phase timing (dist_gpr.TIMER = PhaseTimer() to switch it on)."""

    def __init__(self):
        self.marks = []

    def mark(self, name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.marks.append((name, e))

    def report(self):
        torch.cuda.synchronize()
        acc = {}
        for (n0, e0), (n1, e1) in zip(self.marks[:-1], self.marks[1:]):
            acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1)
        self.marks = []
        return acc


TIMER = None
_FINE = True     # fine-grained marks only make sense when everything runs on one stream
# TRACE = {} switches on a per-panel timeline of the look-ahead schedule: every event the schedule
# records becomes a timing event, and after the factorisation TRACE holds, per event family, the
# milliseconds since the fork of the streams, plus the host time at which each panel was issued
# (tools/dist_trace.py prints it).  Diagnostic only: timing events cost a little on every record.
TRACE = None


def _mark(name, fine=False):
    if TIMER is not None and (_FINE or not fine):
        TIMER.mark(name)


class BlockRowLayout(object):
    """Snake block-cyclic ownership of the block rows of an n x n matrix."""

    def __init__(self, n, block, world):
        if block % 128:
            raise ValueError('block must be a multiple of 128 (the tile of the DMMA kernels)')
        self.n, self.block, self.world = int(n), int(block), int(world)
        self.nblk = (self.n + self.block - 1) // self.block

    def owner(self, b):
        rnd, pos = divmod(b, self.world)
        return pos if rnd % 2 == 0 else self.world - 1 - pos

    def rows(self, b):
        return b * self.block, min(self.n, (b + 1) * self.block)

    def blocks_of(self, rank):
        return [b for b in range(self.nblk) if self.owner(b) == rank]

    def local_offsets(self, rank):
        """block -> first local row, and the number of local rows."""
        off, out = 0, {}
        for b in self.blocks_of(rank):
            r0, r1 = self.rows(b)
            out[b] = off
            off += r1 - r0
        return out, off

    def rows_below(self, rank, k):
        """(first local row, number of local rows) of `rank` in block rows > k."""
        offs, nloc = self.local_offsets(rank)
        lo = nloc
        for b in self.blocks_of(rank):
            if b > k:
                lo = offs[b]
                break
        return lo, nloc - lo

    def inverse_assignment(self):
        """Deal the block rows of K^-1 to ranks, largest (N - c)^2 first, each to the least
        loaded rank (L is replicated, so any assignment is legal).  Returns per-rank sorted
        block lists."""
        load = [0.0] * self.world
        mine = [[] for _ in range(self.world)]
        for b in range(self.nblk):          # cost decreases with b: already largest-first
            r0, r1 = self.rows(b)
            cost = float(self.n - r0) ** 2 * (r1 - r0)
            q = min(range(self.world), key=lambda i: (load[i], i))
            load[q] += cost
            mine[q].append(b)
        return [sorted(m) for m in mine]


# --------------------------------------------------------------------------------- backends
class CudaBackend(object):
    """The product backend: every method is one or two C-ABI calls (no CPU fallback)."""

    def __init__(self, device):
        from . import lib as _L
        self._L = _L
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('the distributed GPR path computes on CUDA devices only')

    def _h(self):
        return self._L.handle_for(self.device)

    poison = False      # tests set this: uninitialised buffers are filled with NaN, so a kernel that
                        # reads what was never written or communicated cannot pass unnoticed

    def empty(self, *shape):
        if self.poison:
            return torch.full(tuple(shape), float('nan'), dtype=F64, device=self.device)
        return torch.empty(*shape, dtype=F64, device=self.device)

    # ---- streams / events for the look-ahead (CUDA streams, no tracing compiler)
    def streams(self):
        """(main, chain, tb, gather, narrow): the caller's stream and four high-priority side streams
        (chain highest: it carries the serial dependency chain of the factorisation)."""
        if getattr(self, '_side', None) is None:
            self._side = (torch.cuda.Stream(self.device, priority=-3),
                          torch.cuda.Stream(self.device, priority=-2),
                          torch.cuda.Stream(self.device, priority=-1),
                          torch.cuda.Stream(self.device, priority=-1))
        return (torch.cuda.current_stream(self.device),) + self._side

    def on(self, stream):
        return torch.cuda.stream(stream)

    def record(self, stream):
        e = torch.cuda.Event(enable_timing=TRACE is not None)
        e.record(stream)
        return e

    def wait(self, stream, event):
        if event is not None:
            stream.wait_event(event)

    def zeros(self, *shape):
        return torch.zeros(*shape, dtype=F64, device=self.device)

    def gram_rows(self, prog, theta, Xr, Xc, out):
        h, L = self._h(), self._L
        vt, vx, vx2, vk = L.view(theta), L.view(Xr), L.view(Xc), L.view(out)
        h.check(h.lib.gps_gram_fwd(h.ptr, ctypes.byref(prog.desc), vt.ref, vx.ref, vx2.ref, 0.0, 0,
                                   vk.ref))

    def potrf_(self, A):
        h, L = self._h(), self._L
        va = L.view(A)
        h.check(h.lib.gps_potrf(h.ptr, va.ref, 0, None))

    def trsm_rlt_(self, Lm, B):
        h, L = self._h(), self._L
        vl, vb = L.view(Lm), L.view(B)
        h.check(h.lib.gps_trsm_rlt(h.ptr, vl.ref, vb.ref))

    def gemm_rowmap_(self, A, B, C, rowlim, coff, flops=-1.0):
        h, L = self._h(), self._L
        va, vb, vc, vr = L.view(A), L.view(B), L.view(C), L.view(rowlim)
        h.check(h.lib.gps_gemm_nt_rowmap(h.ptr, -1.0, va.ref, vb.ref, 1.0, vc.ref, vr.ref, int(coff),
                                         float(flops)))

    # plain device copies of the schedule, behind the backend so that the schedule checker of the
    # tests (tests/test_dist_schedule_cpu.py) sees every access
    def copy_(self, dst, src):
        dst.copy_(src)

    def zero_(self, t):
        t.zero_()

    def unpack_rows_(self, dst, src, index):
        """dst[i, :] = src[index[i], :]"""
        torch.index_select(src, 0, index, out=dst) if dst.is_contiguous() else dst.copy_(src.index_select(0, index))

    def syrk_lower_(self, X, D):
        """D <- D - X X^T, lower triangle only."""
        h, L = self._h(), self._L
        vx, vd = L.view(X), L.view(D)
        h.check(h.lib.gps_gemm_nt(h.ptr, -1.0, vx.ref, vx.ref, 1.0, vd.ref, 0, 0, 1))

    def transpose(self, A):
        h, L = self._h(), self._L
        ld = _round_up(A.shape[0], 16)
        out = torch.empty((A.shape[1], ld), dtype=F64, device=A.device)[:, :A.shape[0]]
        va, vo = L.view(A), L.view(out)
        h.check(h.lib.gps_transpose(h.ptr, va.ref, vo.ref))
        return out

    def transpose_into(self, A, out):
        """out <- A^T (out is a view with its own leading dimension)."""
        h, L = self._h(), self._L
        va, vo = L.view(A), L.view(out)
        h.check(h.lib.gps_transpose(h.ptr, va.ref, vo.ref))

    @staticmethod
    def _starts(row_start, rows):
        import numpy as np
        a = np.ascontiguousarray(row_start, dtype=np.int64)
        assert a.shape == (rows,)
        return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))

    def trsm_rlt_prefix_(self, Lm, B, row_start):
        h, L = self._h(), self._L
        vl, vb = L.view(Lm), L.view(B)
        keep, ptr = self._starts(row_start, B.shape[0])
        h.check(h.lib.gps_trsm_rlt_prefix(h.ptr, vl.ref, vb.ref, ptr))

    def trsm_rln_prefix_(self, Lm, Lt, B, row_start):
        h, L = self._h(), self._L
        vl, vt, vb = L.view(Lm), L.view(Lt), L.view(B)
        keep, ptr = self._starts(row_start, B.shape[0])
        h.check(h.lib.gps_trsm_rln_prefix(h.ptr, vl.ref, vt.ref, vb.ref, ptr))

    def weight_rows_(self, W, grow, beta, block):
        h, L = self._h(), self._L
        vw, vg, vb = L.view(W), L.view(grow), L.view(beta)
        h.check(h.lib.gps_gpr_weight_rows(h.ptr, vw.ref, vg.ref, vb.ref, int(block)))

    def gram_bwd(self, prog, theta, Xr, Xc, W):
        h, L = self._h(), self._L
        dtheta = torch.empty(prog.n_theta, dtype=F64, device=W.device)
        vt, vx, vx2, vw, vd = L.view(theta), L.view(Xr), L.view(Xc), L.view(W), L.view(dtheta)
        h.check(h.lib.gps_gram_bwd(h.ptr, ctypes.byref(prog.desc), vt.ref, vx.ref, vx2.ref, vw.ref,
                                   vd.ref, None))
        return dtheta

    def set_option(self, name, value):
        self._h().set_option(name, int(value))

    def matmul_nt(self, A, B):
        from . import ops
        return ops.gemm_nt(A, B)

    def row_sumsq(self, A):
        from . import ops
        return ops.row_sumsq(A)

    def sum_log_diag(self, Lm):
        h, L = self._h(), self._L
        out = torch.empty(1, dtype=F64, device=Lm.device)
        vl, vo = L.view(Lm), L.view(out)
        h.check(h.lib.gps_sum_log_diag(h.ptr, vl.ref, vo.ref))
        return out[0]


# --------------------------------------------------------------------------------- collectives
class _Comm(object):
    """torch.distributed plumbing (NCCL on GPUs, gloo in the CPU tests); a world of one needs
    no process group at all.  `group` may be a dict {'chain': g1, 'tb': g2, 'gather': g3} of
    process groups over the same ranks: the factorisation issues its three kinds of collectives
    (diagonal-block broadcast, top-block broadcast, panel all-gather) on three communicators so
    that a 100 MB all-gather never sits in front of a 2 MB broadcast of the critical path."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.groups = group if isinstance(group, dict) else {}
        self.group = self.groups.get('gather') if isinstance(group, dict) else group
        if dist.is_available() and dist.is_initialized():
            self.world = dist.get_world_size(self.group)
            self.rank = dist.get_rank(self.group)
        else:
            self.world, self.rank = 1, 0

    def _g(self, which):
        return self.groups.get(which, self.group)

    def global_rank(self, r, g=None):
        g = self.group if g is None else g
        if g is None or self.world == 1:
            return r
        return self.dist.get_global_rank(g, r)

    def broadcast(self, t, src, which='chain'):
        if self.world > 1:
            g = self._g(which)
            self.dist.broadcast(t, src=self.global_rank(src, g), group=g)

    def all_gather(self, out, inp):
        if self.world > 1:
            self.dist.all_gather_into_tensor(out, inp, group=self._g('gather'))
        else:
            out.copy_(inp.reshape(out.shape))

    def all_reduce_sum(self, t):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self._g('gather'))


def _os_environ():
    import os
    return os.environ


# --------------------------------------------------------------------------------- the path
_MAPS = {}


def _layout_maps(lay, P, dev):
    """Static index maps of a layout (cached per shape and device): owner and local row of every
    global row, and for every (panel, rank) the first local row below the panel."""
    key = (lay.n, lay.block, P, str(dev))
    hit = _MAPS.get(key)
    if hit is not None:
        return hit
    own = torch.empty(lay.n, dtype=torch.int64)
    lrow = torch.empty(lay.n, dtype=torch.int64)
    for q in range(P):
        oq, _ = lay.local_offsets(q)
        for b, o in oq.items():
            r0, r1 = lay.rows(b)
            own[r0:r1] = q
            lrow[r0:r1] = torch.arange(o, o + r1 - r0, dtype=torch.int64)
    below_all = [[lay.rows_below(q, k) for q in range(P)] for k in range(lay.nblk)]
    lo_table = torch.tensor([[l for l, _ in row] for row in below_all], dtype=torch.int64)
    if len(_MAPS) > 8:
        _MAPS.clear()
    hit = _MAPS[key] = (own.to(dev), lrow.to(dev), below_all, lo_table.to(dev), {})
    return hit


_FLOPS_CACHE = {}


def factor(prog, theta, noise, X, Yc, lay, comm, be, lookahead=True):
    """Distributed Gram + Cholesky.  Returns (Lfull [N, ld] with the lower block triangle of
    L -- identical on every rank --, Lt [N, ld] = its transpose, alpha_t [R, N] = (L^-1 Yc)^T).

    With `lookahead` the work of panel k is a software pipeline over four CUDA streams; only the
    first carries the serial dependency chain of the factorisation, and it touches nothing but
    512-row objects:
      chain  (highest priority): owner(k) applies the previous panel to its diagonal block (SYRK),
              factors it, broadcasts it; owner(k+1) solves ITS block row k+1 of the panel (the
              "top block") at once.
      tb     top block -> everyone (broadcast), then panel k applied to block column k+1 (the
              block row k+2 of owner(k+2) first: the next top block).
      gather the rest of the panel solve, all-gather of the solved panel -> column k of L (and row
              k of L^T) on every rank, then panel k applied to block column k+2.
      main   panel k applied to block columns >= k+3: the bulk, one masked DMMA GEMM per panel.
    Writers of one block column are ordered main(<= c-3) -> gather(c-2) -> tb(c-1) -> solve(c) by
    events; the chain never waits for an all-gather or a bulk update of the same or the previous
    panel (look-ahead depth 2 with respect to the bulk).
    """
    N, R = Yc.shape
    P, rank = comm.world, comm.rank
    bs = lay.block
    ld = _round_up(N, 16)
    mine = lay.blocks_of(rank)
    offs, nloc = lay.local_offsets(rank)
    dev = X.device

    # ---- local rows of K + noise I, then the ride-along rows (Y - m)^T
    Aloc = be.empty(nloc + R, ld)
    grow = torch.full((nloc + R,), NEVER, dtype=torch.int64, device=dev)
    for b in mine:
        r0, r1 = lay.rows(b)
        o = offs[b]
        be.gram_rows(prog, theta, X[r0:r1], X[:r1], Aloc[o:o + r1 - r0, :r1])
        Aloc[o:o + r1 - r0, r0:r1].diagonal().add_(noise)
        grow[o:o + r1 - r0] = torch.arange(r0, r1, dtype=torch.int64, device=dev)
    Aloc[nloc:, :N] = Yc.t()
    _mark('gram')

    # L (lower block triangle) and L^T (upper), filled column block by column block; what lies
    # outside those triangles is never written and never read
    Lfull = be.empty(N, ld)
    Lt = be.empty(N, ld)
    own_of_row, lrow_of_row, below_all, lo_table, unpack_index = _layout_maps(lay, P, dev)

    Lkk_buf = be.empty(bs * bs)
    top_buf = be.empty(bs * bs)
    mmax_all = max(lay.local_offsets(q)[1] for q in range(P))
    send_buf = be.empty(mmax_all * bs) if P > 1 else None
    recv_buf = be.empty(P * mmax_all * bs) if P > 1 else None

    import numpy as np
    flops_cache = _FLOPS_CACHE.setdefault((lay.n, lay.block, P, rank, R), {})

    def update(k, c_lo, c_hi, Bsrc=None):
        """my rows below block row k, global columns [c_lo, c_hi):  A -= P_k L[c_lo:c_hi, k]^T,
        lower part only (by the global index of each local row); ride-along rows unmasked.
        Bsrc: the rows c_lo:c_hi of panel k when they are not taken from Lfull."""
        if c_hi <= c_lo:
            return
        k0, k1 = lay.rows(k)
        lo, mrows = below_all[k][rank]
        # algorithmic flops of this launch (what the profile option books): a pure function of
        # the layout, so it is computed once per (layout, rank, R) and looked up afterwards --
        # the chain stream's critical path is host-issue bound at 8 ranks, and this loop used to
        # cost ~0.2 ms of host time per panel
        fkey = (k, c_lo, c_hi)
        flops = flops_cache.get(fkey)
        if flops is None:
            flops = 2.0 * (k1 - k0) * R * (c_hi - c_lo)
            for b in mine:
                if b > k:
                    g = np.arange(*lay.rows(b))
                    flops += 2.0 * (k1 - k0) * float(np.clip(np.minimum(g, c_hi - 1) - c_lo + 1, 0, None).sum())
            flops_cache[fkey] = flops
        Bop = Lfull[c_lo:c_hi, k0:k1] if Bsrc is None else Bsrc
        be.gemm_rowmap_(Aloc[lo:, k0:k1], Bop, Aloc[lo:, c_lo:c_hi], grow[lo:], c_lo, flops)

    def factor_diag_and_solve(k):
        """factor the diagonal block, broadcast it, solve my panel rows (+ the ride-along rows)."""
        r0, r1 = lay.rows(k)
        nb = r1 - r0
        own = lay.owner(k)
        Lkk = Lkk_buf[:nb * nb].view(nb, nb)
        if rank == own:
            D = Aloc[offs[k]:offs[k] + nb, r0:r1]
            be.potrf_(D)
            Lkk.copy_(D)
        _mark('potrf_diag', fine=True)
        comm.broadcast(Lkk, own)
        Lfull[r0:r1, r0:r1] = Lkk
        lo, mrows = below_all[k][rank]
        Pn = Aloc[lo:, r0:r1]                       # my panel rows + the ride-along rows
        be.trsm_rlt_(Lkk, Pn)
        _mark('bcast+panel_trsm', fine=True)
        return Pn, mrows

    def gather_panel(k, Pn, mrows):
        """all-gather of the solved panel k -> Lfull[r1:, r0:r1] on every rank; block column k
        of L (diagonal block included) is then also stored transposed, as block row k of Lt."""
        r0, r1 = lay.rows(k)
        nb = r1 - r0
        if k == lay.nblk - 1:
            be.transpose_into(Lfull[r0:r1, r0:r1], Lt[r0:r1, r0:r1])
            return
        if P > 1:
            mmax = max(m for _, m in below_all[k])
            send = send_buf[:mmax * nb].view(mmax, nb)
            if mrows:
                be.copy_(send[:mrows], Pn[:mrows])
            if mrows < mmax:
                be.zero_(send[mrows:])              # padding rows are never unpacked; keep them finite
            recv = recv_buf[:P * mmax * nb]
            comm.all_gather(recv, send.reshape(-1))
            src = unpack_index.get(k)
            if src is None:
                # row of the gathered buffer that holds global row r (static per layout and
                # panel: built once, it saves five small launches per panel on the gather stream)
                oq = own_of_row[r1:]
                src = unpack_index[k] = oq * mmax + lrow_of_row[r1:] - lo_table[k][oq]
            be.unpack_rows_(Lfull[r1:, r0:r1], recv.view(P * mmax, nb), src)
        else:
            be.copy_(Lfull[r1:, r0:r1], Pn[:mrows])
        be.transpose_into(Lfull[r0:, r0:r1], Lt[r0:r1, r0:N])
        _mark('gather+unpack', fine=True)

    _mark('setup')
    streams = be.streams() if lookahead else None
    # The chain of the factorisation runs next to bulk updates that occupy every SM: each extra
    # launch on it waits for an SM to come free, so the split-K slicing of its small products (a
    # gain when the GPU is idle) is switched off for the duration of the factorisation -- the
    # round-1 condition (factor phase 77 ms on 8 GPUs; 87-94 ms in the round-2 A/B runs with it on,
    # profiles/r02_dist_8gpu_ab.txt).  GPSLIM_DIST_SPLITK=1 keeps it on.
    splitk_off = _os_environ().get('GPSLIM_DIST_SPLITK', '0') != '1' and hasattr(be, 'set_option')
    if splitk_off:
        be.set_option('gemm_splitk', 0)
    if streams is None:
        # plain right-looking order
        for k in range(lay.nblk):
            Pn, mrows = factor_diag_and_solve(k)
            gather_panel(k, Pn, mrows)
            if k < lay.nblk - 1:
                update(k, lay.rows(k)[1], N)
                _mark('trailing_gemm', fine=True)
    elif lookahead != 'v2':
        # the round-1 schedule, kept for A/B measurements: chain = factor + broadcast + solve of ALL my
        # panel rows + top-block broadcast + column k+1; gather = all-gather; main = column k+2, rest
        global _FINE
        _FINE = False
        main, chain, _, gath = streams[:4]
        nblk = lay.nblk
        ev_narrow = [None] * nblk
        start = be.record(main)
        be.wait(chain, start)
        be.wait(gath, start)
        ev_c = ev_g = None
        for k in range(nblk):
            with be.on(chain):
                Pn, mrows = factor_diag_and_solve(k)
                ev_trsm = be.record(chain)
                if k + 1 < nblk:
                    r0, r1 = lay.rows(k)
                    t0, t1 = lay.rows(k + 1)
                    nxt = lay.owner(k + 1)
                    T = top_buf[:(t1 - t0) * (r1 - r0)].view(t1 - t0, r1 - r0)
                    if rank == nxt:
                        be.copy_(T, Aloc[offs[k + 1]:offs[k + 1] + t1 - t0, r0:r1])
                    comm.broadcast(T, nxt, 'chain')
                    if k >= 1:
                        be.wait(chain, ev_narrow[k - 1])      # column k+1 has seen panels <= k-1
                    update(k, *lay.rows(k + 1), Bsrc=T)
                ev_c = be.record(chain)
            with be.on(gath):
                be.wait(gath, ev_trsm)
                gather_panel(k, Pn, mrows)
                ev_g = be.record(gath)
            if k + 1 < nblk:
                with be.on(main):
                    be.wait(main, ev_g)
                    if k + 2 < nblk:
                        update(k, *lay.rows(k + 2))
                    ev_narrow[k] = be.record(main)
                    if k + 3 < nblk:
                        update(k, lay.rows(k + 3)[0], N)
        be.wait(main, ev_c)
        be.wait(main, ev_g)
        _FINE = True
        _mark('factor(lookahead)')
    else:
        _FINE = False
        main, chain, tb, gath = streams[:4]
        # the column-(k+2) update gets its own stream: on the gather stream it would hold up the next
        # panel's solve and all-gather while it waits for the bulk update of the previous panel
        import os as _os
        nar = streams[4] if len(streams) > 4 and _os.environ.get('GPSLIM_DIST_NARROW_STREAM', '1') != '0' else gath
        nblk = lay.nblk
        ring = [be.empty(bs * bs) for _ in range(3)]      # diagonal blocks in flight (chain -> gather)
        ev_L, ev_top, ev_colT, ev_col = {}, {}, {}, {}
        ev_solve, ev_g, ev_narrow, ev_rest = {}, {}, {}, {}
        nend = nloc + R
        if TRACE is not None and hasattr(torch.cuda, 'synchronize') and X.is_cuda:
            torch.cuda.synchronize()            # align the host clock with the fork event
        start = be.record(main)
        for st in (chain, tb, gath, nar):
            be.wait(st, start)

        grow_host = np.full(nend, NEVER, dtype=np.int64)
        for b_ in mine:
            g0, g1 = lay.rows(b_)
            grow_host[offs[b_]:offs[b_] + g1 - g0] = np.arange(g0, g1)

        def update_rows(k, a, b, c_lo, c_hi, Bsrc=None):
            """local rows [a, b) of panel k applied to global columns [c_lo, c_hi) (lower part by the
            global index of each row; the ride-along rows unmasked)."""
            if b <= a or c_hi <= c_lo:
                return
            k0, k1 = lay.rows(k)
            fkey = (k, a, b, c_lo, c_hi)
            flops = flops_cache.get(fkey)       # algorithmic flops, a pure function of the layout
            if flops is None:
                g = np.minimum(grow_host[a:b], c_hi - 1)
                flops = flops_cache[fkey] = 2.0 * (k1 - k0) * float(np.clip(g - c_lo + 1, 0, None).sum())
            Bop = Lfull[c_lo:c_hi, k0:k1] if Bsrc is None else Bsrc
            be.gemm_rowmap_(Aloc[a:b, k0:k1], Bop, Aloc[a:b, c_lo:c_hi], grow[a:b], c_lo, flops)

        import time as _time
        host_issue, host_t0 = {}, _time.perf_counter()
        for k in range(nblk):
            if TRACE is not None:
                host_issue[k] = _time.perf_counter()
            r0, r1 = lay.rows(k)
            nb = r1 - r0
            own = lay.owner(k)
            nxt = lay.owner(k + 1) if k + 1 < nblk else -1
            lo_k, m_k = below_all[k][rank]
            Lkk = ring[k % 3][:nb * nb].view(nb, nb)
            if k + 1 < nblk:
                t0, t1 = lay.rows(k + 1)
                nb1 = t1 - t0
                lo1, _ = below_all[k + 1][rank]
            # ------------------------------------------------ chain: the serial dependency chain
            with be.on(chain):
                be.wait(chain, ev_narrow.get(k - 2))      # column k has seen panels <= k-2 ...
                be.wait(chain, ev_rest.get(k - 3))
                if rank == own:
                    D = Aloc[offs[k]:offs[k] + nb, r0:r1]
                    if k >= 1:                            # ... and panel k-1 on the diagonal block, here
                        p0, p1 = lay.rows(k - 1)
                        be.syrk_lower_(Aloc[offs[k]:offs[k] + nb, p0:p1], D)
                    be.potrf_(D)
                    be.copy_(Lkk, D)
                comm.broadcast(Lkk, own, 'chain')
                ev_L[k] = be.record(chain)
                if rank == nxt:
                    be.wait(chain, ev_colT.get(k - 1))    # my block row k+1 has seen panel k-1 in column k
                    be.trsm_rlt_(Lkk, Aloc[offs[k + 1]:offs[k + 1] + nb1, r0:r1])
                    ev_top[k] = be.record(chain)
            # ------------------------------------------------ tb (1): top block -> everyone
            if k + 1 < nblk:
                T = top_buf[:nb1 * nb].view(nb1, nb)
                with be.on(tb):
                    if rank == nxt:
                        be.wait(tb, ev_top[k])
                        be.copy_(T, Aloc[offs[k + 1]:offs[k + 1] + nb1, r0:r1])
                    comm.broadcast(T, nxt, 'tb')
            # ------------------------------------------------ gather (1): the rest of the panel solve
            with be.on(gath):
                be.wait(gath, ev_L[k])
                be.copy_(Lfull[r0:r1, r0:r1], Lkk)
                be.wait(gath, ev_col.get(k - 1))          # column k of my rows has seen panel k-1
                s0 = lo_k + nb1 if rank == nxt else lo_k  # the top block is solved on the chain
                if nend > s0:
                    be.trsm_rlt_(Lkk, Aloc[s0:, r0:r1])
                if rank == nxt:
                    be.wait(gath, ev_top[k])
                ev_solve[k] = be.record(gath)
            # ------------------------------------------------ tb (2): panel k -> block column k+1
            if k + 1 < nblk:
                with be.on(tb):
                    be.wait(tb, ev_solve[k])
                    be.wait(tb, ev_narrow.get(k - 1))     # writers of column k+1: main -> gather -> tb
                    be.wait(tb, ev_rest.get(k - 2))
                    if k + 2 < nblk and rank == lay.owner(k + 2):
                        # block row k+2 first: it is the next top block
                        nb2 = lay.rows(k + 2)[1] - lay.rows(k + 2)[0]
                        update_rows(k, lo1, lo1 + nb2, t0, t1, Bsrc=T)
                        ev_colT[k] = be.record(tb)
                        update_rows(k, lo1 + nb2, nend, t0, t1, Bsrc=T)
                    else:
                        update_rows(k, lo1, nend, t0, t1, Bsrc=T)
                    ev_col[k] = be.record(tb)
            # ------------------------------------------------ gather (2): all-gather, panel k -> column k+2
            with be.on(gath):
                gather_panel(k, Aloc[lo_k:, r0:r1], m_k)
                ev_g[k] = be.record(gath)
            if k + 2 < nblk:
                with be.on(nar):
                    be.wait(nar, ev_g[k])
                    be.wait(nar, ev_rest.get(k - 1))      # main writes columns >= k+2 with panel k-1
                    update_rows(k, lo1, nend, *lay.rows(k + 2))
                    ev_narrow[k] = be.record(nar)
            # ------------------------------------------------ main: the bulk
            if k + 3 < nblk:
                with be.on(main):
                    be.wait(main, ev_g[k])
                    lo3, _ = below_all[k + 2][rank]
                    update_rows(k, lo3, nend, lay.rows(k + 3)[0], N)
                    ev_rest[k] = be.record(main)
        for st in (chain, tb, gath, nar):
            be.wait(main, be.record(st))
        if TRACE is not None and hasattr(start, 'elapsed_time'):
            end = be.record(main)
            end.synchronize()
            fam = dict(L=ev_L, top=ev_top, colT=ev_colT, col=ev_col, solve=ev_solve, gathered=ev_g,
                       narrow=ev_narrow, rest=ev_rest)
            TRACE.clear()
            TRACE.update({name: {k: start.elapsed_time(e) for k, e in d.items()} for name, d in fam.items()})
            TRACE['host_issue_ms'] = {k: (t - host_t0) * 1e3 for k, t in host_issue.items()}
            TRACE['total_ms'] = start.elapsed_time(end)
        _FINE = True
        _mark('factor(lookahead)')
    if splitk_off:
        be.set_option('gemm_splitk', 1)
    alpha_t = Aloc[nloc:, :N]
    return Lfull, Lt, alpha_t


def nlml_and_grad(prog, theta, noise, X, Yc, block=512, group=None, backend=None, want_grad=True,
                  lookahead=True):
    """NLML of GPR and its gradient w.r.t. (theta, noise, Yc), computed by all ranks of
    `group` together.  Inputs are replicated (X is N x D: small); every rank returns the same
    values.  Returns (nlml, dtheta, dnoise, dYc) -- 0-d / [n_theta] / 0-d / [N, R] tensors;
    the gradients are None when want_grad is False."""
    comm = _Comm(group)
    be = backend if backend is not None else CudaBackend(X.device)
    _mark('start')
    N, R = Yc.shape
    if R < 1 or R > 16:
        raise ValueError('1..16 output columns supported')
    noise = float(noise)
    lay = BlockRowLayout(N, block, comm.world)
    Lfull, Ltf, alpha_t = factor(prog, theta, noise, X, Yc, lay, comm, be, lookahead=lookahead)
    Lsq = Lfull[:, :N]
    logdet = be.sum_log_diag(Lsq)
    nlml = 0.5 * N * R * math.log(2.0 * math.pi) + R * logdet + 0.5 * (alpha_t ** 2).sum()
    if not torch.isfinite(nlml):
        from .lib import CholeskyError
        raise CholeskyError('distributed Cholesky failed: K + noise*I is not positive definite')
    if not want_grad:
        return nlml, None, None, None

    out, beta_t = inverse_rows_and_contract(prog, theta, X, lay, comm.rank, Lsq, Ltf[:, :N], alpha_t, be)
    comm.all_reduce_sum(out)
    _mark('contract+allreduce')
    return nlml, out[:prog.n_theta], out[prog.n_theta], beta_t.t().contiguous()


def inverse_rows_and_contract(prog, theta, X, lay, rank, Lsq, Lt, alpha_t, be):
    """This rank's share of the gradient, no communication (L and L^T are replicated): its block
    rows of U = L^-T and of K^-1 = U L^-1 (columns >= the row block), contracted with dK/dtheta.
    Returns (partial [dtheta..., dnoise] to be summed over ranks, beta^T [R, N])."""
    N = Lsq.shape[0]
    R = alpha_t.shape[0]
    dev = X.device
    bs = lay.block
    mine = lay.inverse_assignment()[rank]
    m = R + sum(lay.rows(b)[1] - lay.rows(b)[0] for b in mine)
    ld = _round_up(N, 16)
    B = be.zeros(m, ld)[:, :N]
    B[:R] = alpha_t
    growB = torch.empty(m - R, dtype=torch.int64, device=dev)
    starts, o = [], R
    for b in mine:
        r0, r1 = lay.rows(b)
        B[o:o + r1 - r0, r0:r1].diagonal().fill_(1.0)
        growB[o - R:o - R + r1 - r0] = torch.arange(r0, r1, dtype=torch.int64, device=dev)
        starts.append((r0, o - R, r1 - r0))
        o += r1 - r0
    import numpy as np
    # first column of every row: its block start (alpha rows: 0)
    start_u = np.concatenate([np.full(nr, r0, dtype=np.int64) for (r0, _, nr) in starts]
                             or [np.zeros(0, dtype=np.int64)])
    start_all = np.concatenate([np.zeros(R, dtype=np.int64), start_u])
    _mark('nlml+inv_setup')
    if m > R:
        be.trsm_rlt_prefix_(Lsq, B[R:], start_u)          # rows of U = L^-T
    _mark('rows_of_U')
    be.trsm_rln_prefix_(Lsq, Lt, B, start_all)            # beta^T ; rows of K^-1 (cols >= block)
    _mark('rows_of_Kinv')
    beta_t = B[:R]
    out = torch.zeros(prog.n_theta + 1, dtype=F64, device=dev)
    if m > R:
        W = B[R:]
        be.weight_rows_(W, growB, beta_t, bs)
        out[prog.n_theta] = W[torch.arange(m - R, device=dev), growB].sum()     # tr W
        out[:prog.n_theta] = be.gram_bwd(prog, theta, X.index_select(0, growB), X, W)
    return out, beta_t


def predict(prog, theta, noise, X, Yc, Xnew, kdiag_new, block=512, group=None, backend=None, lookahead=True):
    """Distributed GPR prediction (reference models/gpr.py:118-131): the ranks factor K + noise I
    together (every rank ends up with all of L), then the COLUMNS of K(X, Xnew) -- the test points --
    are dealt out: rank r computes A_r^T = K(Xnew_r, X) L^-T, mean_r = A_r^T V with V = L^-1 (Y - m)
    and var_r = Kdiag(Xnew_r) - sum A_r^2, and one all-gather joins the slices.  Returns
    (mean [N*, R] without the mean function, var [N*]) on every rank."""
    comm = _Comm(group)
    be = backend if backend is not None else CudaBackend(X.device)
    N, R = Yc.shape
    lay = BlockRowLayout(N, block, comm.world)
    Lfull, _, alpha_t = factor(prog, theta, float(noise), X, Yc, lay, comm, be, lookahead=lookahead)
    Lsq = Lfull[:, :N]
    ns = Xnew.shape[0]
    P, rank = comm.world, comm.rank
    per = (ns + P - 1) // P                       # equal slices (the last ones may be short / empty)
    lo, hi = min(ns, rank * per), min(ns, (rank + 1) * per)
    out = be.zeros(per, R + 1)
    if hi > lo:
        At = be.empty(hi - lo, _round_up(N, 16))[:, :N]
        be.gram_rows(prog, theta, Xnew[lo:hi], X, At)          # K(Xnew_r, X) = K(X, Xnew_r)^T
        be.trsm_rlt_(Lsq, At)                                  # (L^-1 Kx)^T
        out[:hi - lo, :R] = be.matmul_nt(At, alpha_t)          # A^T V
        out[:hi - lo, R] = kdiag_new[lo:hi] - be.row_sumsq(At)
    if P > 1:
        full = be.empty(P * per, R + 1)
        comm.all_gather(full.view(-1), out.reshape(-1))
    else:
        full = out
    return full[:ns, :R].contiguous(), full[:ns, R].contiguous()

