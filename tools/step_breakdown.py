"""Per-kernel time of ONE training step from an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py
(steps are delimited by the optimizer's sgd_kernel).  usage: python tools/step_breakdown.py launches.csv [step_index] [-v]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].lstrip('-').isdigit() else 1
verbose = '-v' in sys.argv
rows = list(csv.DictReader([l for l in open(path) if not l.startswith('==')]))
idx = [i for i, r in enumerate(rows) if 'sgd_kernel' in r['Kernel Name']]
step = rows[idx[which] + 1:idx[which + 1] + 1]
tot, cnt = collections.defaultdict(float), collections.Counter()
for i, r in enumerate(step):
    v = float(r['Metric Value'].replace(',', ''))
    u = r['Metric Unit']
    v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
    name = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('b2::', '')
    tot[name] += v
    cnt[name] += 1
    if verbose:
        print("%3d %-44s %-16s %8.1f" % (i, name[:44], r['Grid Size'].replace(' ', ''), v))
print("launches %d   total %.1f us" % (len(step), sum(tot.values())))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
    print("%-64s %4d %9.1f us %5.1f%%" % (k[:64], cnt[k], v, 100 * v / sum(tot.values())))
