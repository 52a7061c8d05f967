"""Hot instruction footprint of a kernel from an .ncu-rep with source counters: distinct SASS instructions whose warp-level
execution count is at least `frac` of the most common loop-level count, x 16 bytes, by function and by source file.
    python tools/ncu_hot_footprint.py gpurun_out/x.ncu-rep [per_iteration_count]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = None
insts = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].startswith("0x"):
        d = dict(zip(hdr, r))
        insts.append(d)
if not insts:
    print("no SASS rows; columns:", rows[:3])
    sys.exit(1)
key = [k for k in hdr if k.startswith("# Warp Instructions Executed") or k == "Instructions Executed"]
kexec = "Instructions Executed" if "Instructions Executed" in hdr else key[0]
ex = [int(d[kexec] or 0) for d in insts]
total = sum(ex)
base = float(sys.argv[2]) if len(sys.argv) > 2 else None
print(f"{len(insts)} SASS instructions ({len(insts) * 16 / 1024:.1f} KB), {total / 1e9:.2f} G warp-instructions executed")
if base is None:
    srt = sorted(ex)
    base = srt[int(len(srt) * 0.5)] or 1
for frac in (1.0, 0.5, 0.1, 0.01):
    hot = [e for e in ex if e >= frac * base]
    print(f"  executed >= {frac:5.2f} x {base:.3g}: {len(hot):6d} instructions = {len(hot) * 16 / 1024:6.1f} KB, "
          f"{100 * sum(hot) / total:5.1f} % of all executed")
