"""ncu launch list (`--metrics gpu__time_duration.sum --csv`) -> share of the run by kernel.
    python profiles/summarize_launches.py gpurun_out/launches.csv "<command that was profiled>" > profiles/rNN_launches_summary.txt"""
import collections
import csv
import sys


def main(path, command):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith('=='))]
    head = rows[0]
    name, value, unit = head.index('Kernel Name'), head.index('Metric Value'), head.index('Metric Unit')
    total, count = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if len(r) <= value:
            continue
        scale = {'ns': 1e-3, 'us': 1., 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1., 'msecond': 1e3}.get(r[unit], 1e-3)
        total[r[name]] += float(r[value].replace(',', '')) * scale
        count[r[name]] += 1
    everything = sum(total.values())
    print('ncu launch list of: %s (cold-cache, serialised)\n' % command)
    for k, v in total.most_common():
        print('%-110s n=%3d total %10.1f us  share %5.1f%%' % (k[:110], count[k], v, 100 * v / everything))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else '?')
