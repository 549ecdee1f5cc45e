"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
Usage: python scripts/launch_summary.py launches.csv [top]"""
import collections
import csv
import sys


def main(path, top=25):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith('==')]
    tot = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        n += 1
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
        d = tot.setdefault(row['Kernel Name'][:80], [0, 0.0])
        d[0] += 1
        d[1] += v
    s = sum(v[1] for v in tot.values())
    print(f'{n} launches, {s / 1e3:.3f} ms total')
    print('| kernel | launches | total ms | share |\n|---|---|---|---|')
    for k, (c, v) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f'| {k} | {c} | {v / 1e3:.3f} | {100 * v / s:.1f}% |')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
