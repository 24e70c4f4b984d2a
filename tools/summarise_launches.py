"""Per-kernel totals of an ncu `--metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys


def main(path):
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
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
        name = d['Kernel Name'].split('(')[0]
        v = float(d['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(d['Metric Unit'], 1.0)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print('%-60s %7s %12s %7s' % ('kernel', 'count', 'total_us', 'share'))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-60s %7d %12.1f %6.1f%%' % (k[:60], v[0], v[1], 100 * v[1] / tot))
    print('%-60s %7d %12.1f' % ('TOTAL', sum(v[0] for v in agg.values()), tot))


if __name__ == '__main__':
    main(sys.argv[1])
