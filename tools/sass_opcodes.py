"""Per-kernel counts of the SASS opcodes that prove the tcgen05 / TMEM / TMA path (cuobjdump -sass of the in-tree library):
UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA load / store), UTCBAR (tcgen05.commit), SYNCS (mbarrier),
HMMA (mma.sync, attention kernels), REDG / ATOMG (global reductions).  python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "hybrid_ctunet_b200/_lib/libctunet_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCCP", "SYNCS", "HMMA", "LDSM", "REDG", "RED", "ATOMG", "ATOM"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("ctu::", "").replace("void ", "")
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        per[cur]["_total"] += 1
        if op in OPS:
            per[cur][op] += 1
        if ".2CTA" in m.group(1):
            per[cur]["2CTA"] += 1
tot = collections.Counter()
print(f"{'kernel':78s} " + " ".join(f"{o:>7s}" for o in OPS) + "   instrs")
for k, c in per.items():
    if any(c[o] for o in OPS):
        print(f"{k[:78]:78s} " + " ".join(f"{c[o]:7d}" for o in OPS) + f"  {c['_total']:7d}")
    for o in OPS:
        tot[o] += c[o]
print(f"{'TOTAL (' + str(len(per)) + ' kernels)':78s} " + " ".join(f"{tot[o]:7d}" for o in OPS))
