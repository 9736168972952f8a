"""Hot spots of one kernel of an .ncu-rep (SASS view): top instructions by warp-stall samples and by excessive
shared-memory wavefronts.  usage: python tools/ncu_hot.py report.ncu-rep kernel-name-regex [n]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
num = lambda r, k: float(r[col[k]]) if r[col[k]] not in ("", "-") else 0.0
tot = sum(num(r, "Warp Stall Sampling (All Samples)") for r in body)
print(f"{len(body)} instructions, {tot:.0f} stall samples, "
      f"{sum(num(r, 'Instructions Executed') for r in body):.0f} warp instructions executed")
print("-- by stall samples")
for i, r in sorted(enumerate(body), key=lambda x: -num(x[1], "Warp Stall Sampling (All Samples)"))[:n]:
    print(f"  {i:5d} {100 * num(r, 'Warp Stall Sampling (All Samples)') / tot:5.1f}%  exec {num(r, 'Instructions Executed'):9.0f}  {r[col['Source']].strip()[:90]}")
print("-- by excessive shared wavefronts")
for i, r in sorted(enumerate(body), key=lambda x: -num(x[1], "L1 Wavefronts Shared Excessive"))[:10]:
    print(f"  {i:5d} excess {num(r, 'L1 Wavefronts Shared Excessive'):9.0f} of {num(r, 'L1 Wavefronts Shared'):9.0f}  {r[col['Source']].strip()[:90]}")
