"""Print the kernels in an .ncu-rep with a few key metrics and the top stall locations.
usage: python tools/ncu_top.py report.ncu-rep [kernel-index]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for ki, r in enumerate(rows[2:]):
    print(f"--- kernel {ki}")
    for k in keys:
        if k in hdr:
            print(f"  {k:70s} {r[hdr.index(k)][:70]} {units[hdr.index(k)]}")
    st = [(float(r[i]), h) for i, h in enumerate(hdr)
          if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    print("  stalls/issue:", ", ".join(f"{h.split('stalled_')[1].split('_per')[0]}={v:.2f}"
                                       for v, h in sorted(st, reverse=True)[:7]))
if len(sys.argv) > 2:
    ki = sys.argv[2]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{ki}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    ia, isrc, ism, iex = (hdr.index(x) for x in ("Address", "Source", "Warp Stall Sampling (All Samples)",
                                                "Instructions Executed"))
    body = [r for r in rows[2:] if len(r) > ism and r[ism].isdigit()]
    tot = sum(int(r[ism]) for r in body)
    by = collections.Counter()
    for r in body:
        toks = r[isrc].split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
        by[op] += int(r[ism])
    print("samples", tot, "; by opcode:", ", ".join(f"{k}={v / tot:.2f}" for k, v in by.most_common(10)))
    for r in sorted(body, key=lambda r: -int(r[ism]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 16]:
        print(f"  {r[ia][-5:]} samples={r[ism]:>5s} exec={r[iex]:>7s}  {r[isrc][:100]}")
