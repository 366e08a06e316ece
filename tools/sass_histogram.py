"""SASS opcode histogram of libb2unet.so (cuobjdump -sass), overall and per kernel for the tensor-core / TMA opcodes."""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "lifelong-nnunet_b200/lib/libb2unet.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
tot, per = collections.Counter(), collections.defaultdict(collections.Counter)
fn = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        tot[m.group(1)] += 1
        per[fn][m.group(1)] += 1
KEY = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "HMMA", "ELECT"]
print("library: %s   instructions: %d   kernels: %d" % (lib, sum(tot.values()), len(per)))
print("tensor-core / TMA / mbarrier opcodes (whole library): " + "  ".join("%s %d" % (k, tot[k]) for k in KEY))
print("\ntop 30 opcodes:")
for k, v in tot.most_common(30):
    print("  %-12s %8d" % (k, v))
print("\nkernels that issue tcgen05.mma (UTCHMMA) / TMA loads (UTMALDG) / TMEM loads (LDTM):")
for f, c in sorted(per.items(), key=lambda kv: -kv[1]["UTCHMMA"]):
    if c["UTCHMMA"] or c["UTMALDG"]:
        print("  %-60s UTCHMMA %4d  UTMALDG %4d  LDTM %3d  UTCBAR %3d  SYNCS %4d" % (f[:60], c["UTCHMMA"], c["UTMALDG"], c["LDTM"], c["UTCBAR"], c["SYNCS"]))
