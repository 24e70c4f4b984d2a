"""Checks the N^3 cost law used to report the reference's CPU path in the workload's unit
(bench.py `cpu_baseline` / `--impl reference`): times the oracle port (torch-CPU fp64 NLML +
autograd gradient, the same function bench.py times) at N_s = 2048 .. 16384 on this machine and
joins the full-size timings of the LAPACK route recorded by oracle/gen_large_golden.py.

    python tools/cpu_oracle_scaling.py > profiles/r02_cpu_oracle_scaling.json
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import oracle_eval  # noqa: E402


def main():
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    sizes = [int(a) for a in sys.argv[1:]] or [2048, 4096, 8192, 16384]
    out = {'what': 'oracle port (torch-CPU fp64, NLML + autograd gradient, D=8), seconds per evaluation',
           'threads': cores, 'autograd_route': {}}
    prev = None
    for n in sizes:
        reps = 3 if n <= 4096 else 2
        t = float(np.min(oracle_eval(n, 8, reps)[1:]))
        row = {'seconds': t, 'seconds_over_n3': t / float(n) ** 3}
        if prev:
            row['ratio_to_previous'] = t / prev[1]
            row['n3_ratio'] = (n / prev[0]) ** 3
        out['autograd_route'][str(n)] = row
        prev = (n, t)
    gl = os.path.join(ROOT, 'tests', 'golden', 'gpr_large_scalars.json')
    if os.path.exists(gl):
        g = json.load(open(gl))
        out['lapack_route'] = {'what': 'oracle/gen_large_golden.py (potrf + potri + blockwise contraction), seconds per '
                                       'evaluation on %s threads' % g.get('threads'),
                               'cases': {k: dict(v['seconds'], factor_and_inverse_over_n3=(v['seconds']['potrf'] + v['seconds']['potri']) / float(k) ** 3)
                                         for k, v in g['cases'].items()}}
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
