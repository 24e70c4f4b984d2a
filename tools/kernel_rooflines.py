"""Per-kernel roofline table for the NON-GEMM kernels of a fused GPR evaluation, derived from an
ncu launch list (`--metrics gpu__time_duration.sum --csv`, see profiles/README.md).

    python tools/kernel_rooflines.py profiles/r01_launches_gpr8192.csv --n 8192 --d 8

For each kernel: launches, mean duration, ALGORITHMIC bytes per launch (what the operation has to
move, from the problem size -- or from the launch grid for the transposes), achieved GB/s and
its fraction of the HBM peak.  The times are ncu's serialised cold-cache per-launch times, so
they bound the in-pipeline rates from below.  HBM peak: MEASURED_PEAKS.json if the driver wrote
one, else the fallback stated in /opt/skills/guides/B200_PROFILING.md (6.65 TB/s, "of fallback").
"""
import argparse
import collections
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def hbm_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        for k, v in d.items():
            if 'hbm' in k.lower() and isinstance(v, (int, float)):
                return float(v) * (1e3 if v < 100 else 1.0), 'MEASURED_PEAKS.json:' + k
    return 6650.0, 'fallback (B200_PROFILING.md)'


def rows(path):
    hdr = None
    for r in csv.reader(open(path)):
        if len(r) < 6:
            continue
        if r[0] == 'ID':
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        if d.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        us = float(d['Metric Value'].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}[d['Metric Unit']]
        grid = tuple(int(v) for v in re.findall(r'\d+', d['Grid Size']))
        yield d['Kernel Name'].split('(')[0], grid, us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('csv')
    ap.add_argument('--n', type=int, required=True)
    ap.add_argument('--d', type=int, default=8)
    ap.add_argument('--r', type=int, default=1)
    a = ap.parse_args()
    N, D, R = a.n, a.d, a.r
    tri = 8.0 * N * (N + 1) / 2                      # one triangle of an N x N fp64 matrix
    model = {                                        # kernel substring -> (bytes per launch, what)
        'gemv_kernel': (tri + 16.0 * N, 'beta = U alpha: upper triangle of U read once'),
        'gram_fwd_stat_kernel': (tri + 8.0 * N * (D + 1), 'lower triangle of K written once'),
        'gram_fwd_kernel': (tri + 8.0 * N * (D + 1), 'lower triangle of K written once (interpreter)'),
        'gram_bwd_stat_kernel': (tri + 8.0 * N * (D + 1 + R), 'lower triangle of K^-1 read once'),
        'gram_bwd_kernel': (tri + 8.0 * N * (D + 1 + R), 'lower triangle of K^-1 read once (interpreter)'),
    }
    agg = collections.OrderedDict()
    for name, grid, us in rows(a.csv):
        short = name.split('::')[-1].split('<')[0].strip()
        if short == 'transpose_kernel':
            b = 16.0 * 1024 * grid[0] * grid[1]      # 32 x 32 tile read + written per CTA
            what = 'tile read + written (bytes from the launch grid)'
        elif short in model:
            b, what = model[short]
        elif short in ('potrf_leaf_kernel', 'trsm_strip_kernel', 'feature_kernel', 'reduce_cols_kernel',
                       'nlml_kernel', 'u_diag_kernel', 'put_yt_kernel'):
            b, what = None, 'latency-bound / tiny'
        else:
            continue
        e = agg.setdefault(short, [0, 0.0, 0.0, what])
        e[0] += 1
        e[1] += us
        e[2] += b or 0.0
    peak, src = hbm_peak()
    print('workload: fused GPR evaluation N=%d D=%d R=%d   HBM peak %.0f GB/s (%s)' % (N, D, R, peak, src))
    print('%-22s %6s %10s %12s %9s %6s  %s' % ('kernel', 'count', 'mean_us', 'MB/launch', 'GB/s', 'frac', 'algorithmic traffic'))
    for k, (cnt, us, byt, what) in agg.items():
        if byt:
            gbs = byt / us * 1e-3
            print('%-22s %6d %10.1f %12.1f %9.0f %6.2f  %s' % (k, cnt, us / cnt, byt / cnt / 1e6, gbs, gbs / peak, what))
        else:
            print('%-22s %6d %10.1f %12s %9s %6s  %s' % (k, cnt, us / cnt, '-', '-', '-', what))


if __name__ == '__main__':
    main()
