"""Race check of the look-ahead schedules (v1: three streams, the default; v2: five streams) of the distributed factorisation
(gpflowSlim/_backend/dist_gpr.py:factor) WITHOUT a GPU: a tracing backend records, for every
operation the schedule issues, its stream and the memory regions it reads and writes, plus every
event record / wait; vector clocks over the streams then say which pairs of operations are
ordered.  Every pair of conflicting accesses (write-write or read-write on overlapping regions)
issued on different streams must be ordered by events -- exactly what CUDA guarantees and what
the issue-order execution of the other CPU tests cannot see."""
import contextlib

import numpy as np
import pytest
import torch

from gpflowSlim._backend import dist_gpr

F64 = torch.float64


class Region(object):
    __slots__ = ('base', 'r0', 'r1', 'c0', 'c1')

    def __init__(self, t, ld_big):
        self.base = t.untyped_storage().data_ptr()
        off = t.storage_offset()
        if t.dim() == 2 and t.shape[0] >= 1 and t.stride(0) == ld_big and t.shape[1] <= ld_big:
            self.r0, self.c0 = divmod(off, ld_big)
            self.r1, self.c1 = self.r0 + t.shape[0], self.c0 + t.shape[1]
        else:                                   # flat scratch buffer: one "row" of elements
            assert t.is_contiguous(), (t.shape, t.stride())
            self.r0, self.r1, self.c0, self.c1 = 0, 1, off, off + t.numel()

    def overlaps(self, o):
        return (self.base == o.base and self.r0 < o.r1 and o.r0 < self.r1 and self.c0 < o.c1 and o.c0 < self.c1)


class Tracer(object):
    """Backend + communicator stand-in that computes nothing and logs everything."""

    def __init__(self, ld_big, world, rank):
        self.ld, self.world, self.rank = ld_big, world, rank
        self.cur = 'main'
        self.clock = {}          # stream -> vector clock (dict stream -> count)
        self.ops = []            # (name, stream, vc, reads, writes)

    # ---- streams / events
    def streams(self):
        return 'main', 'chain', 'tb', 'gather', 'narrow'

    @contextlib.contextmanager
    def on(self, stream):
        old, self.cur = self.cur, stream
        try:
            yield
        finally:
            self.cur = old

    def _vc(self, s):
        return self.clock.setdefault(s, {})

    def record(self, stream):
        return dict(self._vc(stream))

    def wait(self, stream, event):
        if event is None:
            return
        vc = self._vc(stream)
        for k, v in event.items():
            if vc.get(k, 0) < v:
                vc[k] = v

    def _op(self, name, reads, writes):
        s = self.cur
        vc = self._vc(s)
        vc[s] = vc.get(s, 0) + 1
        self.ops.append((name, s, dict(vc), [Region(t, self.ld) for t in reads if t is not None and t.numel()],
                         [Region(t, self.ld) for t in writes if t is not None and t.numel()]))

    # ---- memory
    def empty(self, *shape):
        return torch.zeros(tuple(shape), dtype=F64)

    zeros = empty

    # ---- kernels
    def gram_rows(self, prog, theta, Xr, Xc, out):
        pass                                     # before the streams fork

    def potrf_(self, A):
        self._op('potrf', [A], [A])

    def trsm_rlt_(self, Lm, B):
        self._op('trsm', [Lm, B], [B])

    def syrk_lower_(self, X, D):
        self._op('syrk', [X, D], [D])

    def gemm_rowmap_(self, A, B, C, rowlim, coff, flops=-1.0):
        self._op('gemm', [A, B, C], [C])

    def transpose_into(self, A, out):
        self._op('transpose', [A], [out])

    def copy_(self, dst, src):
        self._op('copy', [src], [dst])

    def zero_(self, t):
        self._op('zero', [], [t])

    def unpack_rows_(self, dst, src, index):
        self._op('unpack', [src], [dst])

    # ---- collectives (synchronised with the current stream on both sides)
    def broadcast(self, t, src, which='chain'):
        self._op('bcast_' + which, [t] if src == self.rank else [], [] if src == self.rank else [t])

    def all_gather(self, out, inp):
        self._op('all_gather', [inp], [out])

    def all_reduce_sum(self, t):
        self._op('all_reduce', [t], [t])


def _happens_before(a, b):
    """op a (issued earlier) is ordered before op b by stream order / events."""
    sa = a[1]
    return a[2][sa] <= b[2].get(sa, 0)


@pytest.mark.parametrize('schedule', ['v1', 'v2'])
@pytest.mark.parametrize('world,nblk', [(1, 7), (2, 9), (3, 10), (4, 13), (8, 20)])
def test_lookahead_schedule_has_no_unordered_conflicts(world, nblk, schedule):
    bs, R = 128, 2
    N = nblk * bs - 37                    # ragged last block
    ld = dist_gpr._round_up(N, 16)
    lay = dist_gpr.BlockRowLayout(N, bs, world)

    class Prog(object):
        n_theta = 3
    for rank in range(world):
        tr = Tracer(ld, world, rank)
        X = torch.zeros(N, 2, dtype=F64)
        Yc = torch.zeros(N, R, dtype=F64)
        Lfull, Lt, alpha_t = dist_gpr.factor(Prog(), torch.zeros(3, dtype=F64), 0.1, X, Yc, lay, tr, tr,
                                             lookahead=schedule)
        # the consumer on the main stream reads everything
        tr._op('consume', [Lfull[:, :N], Lt[:, :N], alpha_t], [])
        ops = tr.ops
        bad = []
        for j, b in enumerate(ops):
            for i in range(j):
                a = ops[i]
                if a[1] == b[1]:
                    continue
                conflict = (any(w.overlaps(x) for w in a[4] for x in b[3] + b[4]) or
                            any(r.overlaps(w) for r in a[3] for w in b[4]))
                if conflict and not _happens_before(a, b):
                    bad.append((i, a[0], a[1], j, b[0], b[1]))
        assert not bad, 'rank %d: %d unordered conflicting pairs, first: %s' % (rank, len(bad), bad[:5])
        # the schedule really is concurrent: most cross-stream pairs are NOT ordered
        streams_used = {o[1] for o in ops}
        want = {'main', 'chain', 'tb', 'gather', 'narrow'} if schedule == 'v2' else {'main', 'chain', 'gather'}
        assert streams_used == want or nblk < 4


def test_the_checker_sees_a_missing_wait(monkeypatch):
    """Sanity of the checker itself: drop the event waits of one stream and conflicts must show."""
    world, nblk, bs, R = 2, 8, 128, 1
    N = nblk * bs
    ld = dist_gpr._round_up(N, 16)
    lay = dist_gpr.BlockRowLayout(N, bs, world)

    class Prog(object):
        n_theta = 3

    class Sloppy(Tracer):
        def wait(self, stream, event):
            if stream == 'tb':               # the tb stream ignores every dependency
                return
            Tracer.wait(self, stream, event)
    tr = Sloppy(ld, world, 0)
    dist_gpr.factor(Prog(), torch.zeros(3, dtype=F64), 0.1, torch.zeros(N, 2, dtype=F64), torch.zeros(N, R, dtype=F64),
                    lay, tr, tr, lookahead='v2')
    ops = tr.ops
    n_bad = 0
    for j, b in enumerate(ops):
        for i in range(j):
            a = ops[i]
            if a[1] != b[1] and (any(w.overlaps(x) for w in a[4] for x in b[3] + b[4]) or
                                 any(r.overlaps(w) for r in a[3] for w in b[4])) and not _happens_before(a, b):
                n_bad += 1
    assert n_bad > 0
