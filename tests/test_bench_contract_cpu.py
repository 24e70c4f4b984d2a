"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the keys
the driver reads, ranks other than 0 stay silent, the product arm refuses to run without CUDA, and
the block-row layout handles ragged sizes."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True,
                          text=True, timeout=300, env=e)


def test_reference_arm_prints_the_contract_line():
    out = _run(['--impl', 'reference', '--steps', '1', '--warmup', '0', '--cpu-n', '256', '--size', '1024'])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'evals/s' and d['higher_is_better'] is True
    assert d['metric'] == 'GPR NLML+grad evals/s' and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert 'N_s=256' in d['cpu_baseline']['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 'evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['config']['workload'].startswith('GPR ARD-RBF N=1024')


def test_reference_arm_other_ranks_do_no_work():
    out = _run(['--impl', 'reference', '--steps', '1', '--warmup', '0', '--cpu-n', '256'], env={'RANK': '1'})
    assert out.returncode == 0 and out.stdout.strip() == ''


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the behaviour on a machine without CUDA')
def test_product_arm_refuses_to_run_without_cuda():
    out = _run(['--steps', '1', '--warmup', '0'])
    assert out.returncode != 0
    assert 'no CPU fallback' in (out.stderr + out.stdout)


@pytest.mark.parametrize('n,block,world', [(1000, 256, 3), (130, 128, 2), (4096, 512, 8), (777, 128, 1)])
def test_block_row_layout_ragged(n, block, world):
    sys.path.insert(0, os.path.join(ROOT, 'gpflow-slim_b200'))
    from gpflowSlim._backend.dist_gpr import BlockRowLayout
    lay = BlockRowLayout(n, block, world)
    rows = []
    for q in range(world):
        offs, nloc = lay.local_offsets(q)
        assert nloc == sum(lay.rows(b)[1] - lay.rows(b)[0] for b in lay.blocks_of(q))
        for b in lay.blocks_of(q):
            r0, r1 = lay.rows(b)
            assert 0 <= r0 < r1 <= n and lay.owner(b) == q
            rows += list(range(r0, r1))
        for k in range(lay.nblk):
            lo, m = lay.rows_below(q, k)
            assert m == sum(lay.rows(b)[1] - lay.rows(b)[0] for b in lay.blocks_of(q) if b > k)
            assert lo + m == nloc
    assert sorted(rows) == list(range(n))
    inv = lay.inverse_assignment()
    assert sorted(b for m in inv for b in m) == list(range(lay.nblk))
    with pytest.raises(ValueError):
        BlockRowLayout(n, 100, world)
