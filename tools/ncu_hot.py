"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv` (one kernel per file section).
usage: ncu -i rep --page source --csv > src.csv; python tools/ncu_hot.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = None
sect = []
for r in rows:
    if r and r[0] == "Kernel Name":
        print("==", r[1][:110])
        continue
    if r and r[0] == "Address":
        hdr = r
        sect = []
        continue
    if hdr and len(r) == len(hdr):
        sect.append(r)
if hdr:
    si = hdr.index("# Samples")
    ii = hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[si] or 0) for r in sect)
    print("total samples", tot)
    for idx, r in sorted(enumerate(sect), key=lambda t: -int(t[1][si] or 0))[:top]:
        st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        print("%5d %5.1f%%  #%4d  exec %8s  %-70s %s" % (int(r[si] or 0), 100.0 * int(r[si] or 0) / max(tot, 1), idx, r[ii], r[1].strip()[:70],
                                                     " ".join("%s:%d" % (n, c) for c, n in st if c)))
