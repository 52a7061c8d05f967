"""Hot instruction bytes per device function of one kernel: joins the per-instruction execution counts of an .ncu-rep
(source page, SASS) with the function layout inside the kernel's .text (nvdisasm labels of the object file).
    python tools/ncu_hot_by_function.py rep.ncu-rep obj.o kernel-substring per_iteration_count"""
import csv, glob, os, re, subprocess, sys, tempfile, collections
rep, obj, want, base = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3], float(sys.argv[4])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr, insts = None, []
for r in rows:
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].startswith("0x"):
        insts.append((int(r[0], 16), int(r[hdr.index("Instructions Executed")] or 0), r[1].strip()))
a0 = insts[0][0]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, capture_output=True)
out = subprocess.run(["nvdisasm", "-c", glob.glob(os.path.join(d, "*.cubin"))[0]], capture_output=True, text=True).stdout
kern, marks, last, found = None, [], 0, None
for ln in out.splitlines():
    m = re.match(r'\.text\.(\S+):', ln)
    if m:
        if kern and want in kern and found is None: found = marks + [("<end>", last + 16)]
        kern, marks, last = m.group(1), [("<main body>", 0)], 0; continue
    m = re.match(r'(\$\S+):', ln)
    if m and kern: marks.append((m.group(1).split('$')[-1], last + 16)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/', ln)
    if m: last = int(m.group(1), 16)
if kern and want in kern and found is None: found = marks + [("<end>", last + 16)]
names = [re.sub(r'^_ZN\d+_INTERNAL_[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+', '', n)[:60] for n, _ in found]
starts = [s for _, s in found]
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
import bisect
for addr, ex, txt in insts:
    k = bisect.bisect_right(starts, addr - a0) - 1
    a = agg[names[k]]
    a[0] += 1; a[1] += ex
    if ex >= base: a[2] += 1
    if ex >= 0.5 * base: a[3] += 1
print(f"{'function':62s} {'KB':>6s} {'KB >=1/it':>10s} {'KB >=.5/it':>10s} {'M instr':>9s}")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][3]):
    print(f"{n:62s} {a[0] * 16 / 1024:6.1f} {a[2] * 16 / 1024:10.1f} {a[3] * 16 / 1024:10.1f} {a[1] / 1e6:9.0f}")
