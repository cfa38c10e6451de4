"""Per-launch table (duration, active lanes, warp instructions) from an ncu csv with several metrics per launch.
    python profiles/launch_table.py gpurun_out/launches.csv [min_us]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith('=='))]
floor = float(sys.argv[2]) if len(sys.argv) > 2 else 150.
h = rows[0]
ix = {n: i for i, n in enumerate(h)}
per = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= ix['Metric Value']:
        continue
    key = (r[ix['ID']], r[ix['Kernel Name']][:70])
    per.setdefault(key, {})[r[ix['Metric Name']]] = float(r[ix['Metric Value']].replace(',', ''))
for (i, k), m in per.items():
    us = m.get('gpu__time_duration.sum', 0) / 1e3
    if us > floor:
        print('%4s %-70s %8.2f ms  lanes %4.1f  warp inst %.3g' % (
            i, k, us / 1e3, m.get('smsp__thread_inst_executed_per_inst_executed.ratio', 0),
            m.get('smsp__inst_executed.sum', 0)))
