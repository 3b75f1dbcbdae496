"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys


def main(path, top=25):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row['Metric Value'].replace(',', ''))
        except (KeyError, ValueError):
            continue
        unit = row.get('Metric Unit', 'us')
        v = v / 1e3 if unit == 'ns' else v * 1e3 if unit == 'ms' else v
        a = agg[row['Kernel Name'][:90]]
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f'# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.1f} us total (cold-cache, serialised)')
    print(f'{"us":>10} {"n":>5} {"share":>6}  kernel')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f'{v[1]:10.1f} {v[0]:5d} {100 * v[1] / tot:5.1f}%  {k}')


if __name__ == '__main__':
    main(sys.argv[1])
