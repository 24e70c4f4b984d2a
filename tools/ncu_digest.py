"""Digest of an `ncu --set full` report: per captured kernel, the handful of numbers this project
argues with (duration, registers / shared memory / occupancy, DRAM bytes, L2 hit rate, the busiest
pipes, shared-memory bank conflicts, and the warp-stall reasons per issued instruction, sorted).

    python tools/ncu_digest.py gpurun_out/x.ncu-rep [--json out.json] [-k regex]

Reads the report with `ncu -i <rep> --page raw --csv` (works without a GPU) or takes that CSV
directly.  Written at the end of round 1 for the round-2 kernel work (leaf Cholesky, prefix
solves); checked against profiles/r01_gemm8192_tma_full.ncu-rep."""
import argparse
import csv
import io
import json
import re
import subprocess
import sys


def raw_rows(path):
    if path.endswith('.csv'):
        text = open(path).read()
    else:
        text = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True,
                              check=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    while rows and (not rows[0] or rows[0][0] != 'ID'):
        rows.pop(0)
    return rows[0], rows[1], rows[2:]


def num(s):
    try:
        return float(s.replace(',', ''))
    except (ValueError, AttributeError):
        return None


SCALE = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 's': 1.0,
         'msecond': 1e-3, 'usecond': 1e-6, 'nsecond': 1e-9, 'second': 1.0}


def digest(hdr, units, row):
    col = {h: i for i, h in enumerate(hdr)}

    def get(name, scaled=False):
        i = col.get(name)
        if i is None:
            return None
        v = num(row[i])
        if v is not None and scaled:
            v *= SCALE.get(units[i].split('/')[0], 1.0)
        return v
    d = {'kernel': row[col['Kernel Name']] if 'Kernel Name' in col else '?',
         'grid': row[col['Grid Size']] if 'Grid Size' in col else None,
         'block': row[col['Block Size']] if 'Block Size' in col else None,
         'duration_ms': (get('gpu__time_duration.sum', True) or 0.0) * 1e3,
         'registers_per_thread': get('launch__registers_per_thread'),
         'dyn_smem_bytes': get('launch__shared_mem_per_block_dynamic', True),
         'static_smem_bytes': get('launch__shared_mem_per_block_static', True),
         'achieved_occupancy_pct': get('sm__warps_active.avg.pct_of_peak_sustained_active'),
         'dram_read_bytes': get('dram__bytes_read.sum', True),
         'dram_write_bytes': get('dram__bytes_write.sum', True),
         'l2_hit_pct': get('lts__t_sector_hit_rate.pct'),
         'sm_throughput_pct': get('sm__throughput.avg.pct_of_peak_sustained_elapsed'),
         'smem_bank_conflicts': get('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'),
         'warp_latency_per_inst': get('smsp__average_warp_latency_per_inst_issued.ratio')}
    if d['dram_read_bytes'] is not None and d['duration_ms']:
        d['dram_GBps'] = (d['dram_read_bytes'] + (d['dram_write_bytes'] or 0.0)) / d['duration_ms'] / 1e6
    pipes = {}
    for h, i in col.items():
        m = re.match(r'sm__inst_executed_pipe_(\w+)\.avg\.pct_of_peak_sustained_active$', h) or \
            re.match(r'sm__pipe_(\w+)_cycles_active\.avg\.pct_of_peak_sustained_active$', h)
        if m and num(row[i]) is not None:
            pipes[m.group(1)] = max(pipes.get(m.group(1), 0.0), num(row[i]))
    d['pipes_pct_of_peak_active'] = dict(sorted(pipes.items(), key=lambda kv: -kv[1])[:6])
    stalls = {}
    for h, i in col.items():
        m = re.match(r'smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio$', h)
        if m and num(row[i]):
            stalls[m.group(1)] = num(row[i])
    d['stalls_per_issue'] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('-k', default=None, help='regex on the kernel name')
    ap.add_argument('--json', default=None)
    a = ap.parse_args()
    hdr, units, rows = raw_rows(a.report)
    out = []
    for row in rows:
        if len(row) != len(hdr):
            continue
        d = digest(hdr, units, row)
        if a.k and not re.search(a.k, d['kernel']):
            continue
        out.append(d)
        print('%s  grid %s block %s' % (d['kernel'][:70], d['grid'], d['block']))
        print('  %.3f ms   regs %s   smem %s+%s B   occupancy %s %%   SM throughput %s %%' % (
            d['duration_ms'], d['registers_per_thread'], d['dyn_smem_bytes'], d['static_smem_bytes'],
            d['achieved_occupancy_pct'], d['sm_throughput_pct']))
        if d.get('dram_GBps') is not None:
            print('  DRAM %.3f GB read + %.3f GB written = %.0f GB/s   L2 hit %s %%   smem bank conflicts %s' % (
                d['dram_read_bytes'] / 1e9, (d['dram_write_bytes'] or 0) / 1e9, d['dram_GBps'], d['l2_hit_pct'],
                d['smem_bank_conflicts']))
        print('  pipes (%% of peak, active): %s' % ', '.join('%s %.1f' % kv for kv in d['pipes_pct_of_peak_active'].items()))
        print('  stalls per issued instruction: %s' % ', '.join('%s %.2f' % kv for kv in d['stalls_per_issue'].items()))
    if a.json:
        json.dump(out, open(a.json, 'w'), indent=1)
    return 0


if __name__ == '__main__':
    sys.exit(main())
