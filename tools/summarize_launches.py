"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import collections
import csv
import re
import sys

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lines = [l for l in open(path) if not l.startswith('==')]
tot, cnt = collections.defaultdict(float), collections.Counter()
rows = list(csv.DictReader(lines))
for row in rows[skip:]:
    try:
        v = float(row['Metric Value'].replace(',', ''))
    except ValueError:
        continue
    u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    name = re.sub(r'^void ', '', name)
    tot[name] += v
    cnt[name] += 1
allt = sum(tot.values())
print("launches %d   total %.1f us" % (sum(cnt.values()), allt))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:30]:
    print("%-72s %6d %10.1f us %5.1f%%" % (k[:72], cnt[k], v, 100 * v / allt))
