"""Summarise an .ncu-rep (read here, no GPU): key raw metrics + stall samples aggregated by source line.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [top_n]
"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
print("== raw metrics (%s)" % rep)
for h, u, v in zip(hdr, units, vals):
    if h in want or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        print(f"{h:90s} {u:12s} {v}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur = None
agg, inst, text = collections.Counter(), collections.Counter(), {}
tot = 0
n_sass = 0
for r in csv.reader(src.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) >= 8 and r[0] == "" and r[2].startswith("0x"):
        n_sass += 1
    if len(r) < 8 or r[0] in ("Line No", ""):
        continue
    try:
        ln, s, ie = int(r[0]), int(r[6]), int(r[7])
    except ValueError:
        continue
    agg[(cur, ln)] += s
    inst[(cur, ln)] += ie
    text[(cur, ln)] = r[1]
    tot += s
print("\n== stall samples by file (total %d)" % tot)
byfile = collections.Counter()
for (f, l), s in agg.items():
    byfile[f] += s
for f, s in byfile.most_common():
    print(f"{f:32s} {100 * s / tot:5.1f}%")
print("\n== top source lines: file:line  samples%  warp-instructions(M)  source")
for (f, l), s in agg.most_common(top_n):
    print(f"{f}:{l:<4d} {100 * s / tot:5.1f}% {inst[(f, l)] / 1e6:9.0f}M  {text[(f, l)].strip()[:100]}")
