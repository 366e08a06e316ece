"""per-kernel-class totals of one training step from ncu launch lists (gpu__time_duration), side by side"""
import collections, csv, re, sys
def load(path):
    rows=list(csv.DictReader([l for l in open(path) if not l.startswith('==')]))
    starts=[i for i,r in enumerate(rows) if 'nchw_to_ndhwc' in r['Kernel Name']]
    s,e=starts[-2],starts[-1]
    return rows[s:e]
def name(r):
    n=re.sub(r'\(.*','',r['Kernel Name']).replace('void ','').replace('b2::','')
    return n[:48]
files=sys.argv[1:]
tabs=[]
for f in files:
    t=collections.defaultdict(lambda:[0,0.0])
    for r in load(f):
        v=float(r['Metric Value'])/1e3
        t[name(r)][0]+=1; t[name(r)][1]+=v
    tabs.append(t)
keys=sorted(set().union(*[set(t) for t in tabs]), key=lambda k:-max(t[k][1] if k in t else 0 for t in tabs))
print("%-50s"%"kernel"+"".join("%18s"%f.split('/')[-1][:16] for f in files))
for k in keys:
    print("%-50s"%k+"".join("%6d %9.1f  "%(t[k][0],t[k][1]) if k in t else "%18s"%"-" for t in tabs))
print("%-50s"%"TOTAL"+"".join("%6d %9.1f  "%(sum(v[0] for v in t.values()),sum(v[1] for v in t.values())) for t in tabs))
