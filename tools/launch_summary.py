"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import io
import re
import sys

src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
lines = [l for l in open(src) if not l.startswith("==")]
tot = collections.defaultdict(lambda: [0, 0.0])
rows = list(csv.DictReader(io.StringIO("".join(lines))))
# optional 5th argument "step=K": only the launches of ONE train step -- those after the K-th adam_kernel launch of the
# list up to and including the (K+1)-th (every step ends with the optimizer's single launch)
step = next((int(a[5:]) for a in sys.argv[4:] if a.startswith("step=")), None)
if step is not None:
    ends = [i for i, r in enumerate(rows) if "adam_kernel" in r["Kernel Name"]]
    rows = rows[ends[step - 1] + 1:ends[step] + 1]
for row in rows:
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:90]
    tot[name][0] += 1
    tot[name][1] += v
T = sum(v[1] for v in tot.values())
out = [f"# {title}", f"# total kernel time {T / 1e3:.3f} ms over {sum(v[0] for v in tot.values())} launches "
       "(gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)",
       "kernel,launches,total_us,share"]
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{k},{v[0]},{v[1]:.1f},{v[1] / T:.3f}")
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[:int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4].isdigit() else 25]))
