"""Per-kernel counts of the SASS mnemonics that identify the tensor-core / TMA / cluster paths, from the in-tree
library.  usage: python tools/sass_summary.py [libsparch_b200.so] > profiles/rNN_sass_summary.txt

UTCHMMA / UTCIMMA = tcgen05.mma kind::f16 / kind::i8, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st,
UTMALDG / UTMASTG = TMA tensor loads / stores, HMMA / IMMA = legacy mma.sync, STAS = st.async (DSMEM),
SYNCS = mbarrier operations, UCGABAR = cluster barrier."""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        "sparch_b200", "libsparch_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "HMMA", "IMMA", "STAS", "SYNCS", "UCGABAR",
        "LDG.E.128.STRONG", "STG.E.STRONG", "MEMBAR.ALL.GPU", "STL", "LDL"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name)
        counts[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in line:
        continue
    ins = line.split("*/", 1)[1] if "*/" in line else ""
    total[cur] += bool(re.match(r"\s+(@!?U?P\d\s+)?[A-Z]", ins))
    for k in keys:
        if re.search(r"\b" + re.escape(k), ins):
            counts[cur][k] += 1
print(f"# SASS summary of {os.path.basename(lib)} (cuobjdump -sass, sm_100a): instruction counts per kernel")
print("kernel,instructions," + ",".join(keys))
for k, c in counts.items():
    if any(c.values()):
        print(f"{k},{total[k]}," + ",".join(str(c[x]) for x in keys))
print("# kernels without any of these mnemonics:", ", ".join(k for k, c in counts.items() if not any(c.values())))
