"""Hot instruction bytes per SOURCE LINE of one kernel: per-instruction execution counts of an .ncu-rep (source page, SASS)
joined by offset with nvdisasm -g line markers of the SAME build's object file.
    python tools/ncu_hot_by_line.py rep.ncu-rep obj.o kernel-substring per_iteration_count [top_n]"""
import bisect, collections, csv, glob, os, re, subprocess, sys, tempfile
rep, obj, want, base = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3], float(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 60
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
hdr, insts = None, []
for r in csv.reader(src.splitlines()):
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].startswith("0x"):
        insts.append((int(r[0], 16), int(r[hdr.index("Instructions Executed")] or 0)))
a0 = insts[0][0]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, capture_output=True)
out = subprocess.run(["nvdisasm", "-g", "-c", glob.glob(os.path.join(d, "*.cubin"))[0]], capture_output=True, text=True).stdout
kern, cur, fn, where = None, None, "<main>", {}
for ln in out.splitlines():
    m = re.match(r'\.text\.(\S+):', ln)
    if m: kern = m.group(1); fn = "<main>"; continue
    if not kern or want not in kern: continue
    m = re.match(r'(\$\S+):', ln)
    if m: fn = re.sub(r'^.*_cu_[0-9a-f]{8}', '', m.group(1).split('$')[-1])[:28]; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/', ln)
    if m: where[int(m.group(1), 16)] = (fn, cur)
agg = collections.defaultdict(lambda: [0, 0, 0])
byfn = collections.defaultdict(lambda: [0, 0, 0])
miss = 0
for addr, ex in insts:
    w = where.get(addr - a0)
    if w is None: miss += 1; continue
    for tab, key in ((agg, w), (byfn, w[0])):
        a = tab[key]; a[0] += 16; a[2] += ex
        if ex >= 0.5 * base: a[1] += 16
print(f"{len(insts)} instructions, {miss} unmatched; hot = executed by some warp in >= half of the iterations")
print("== hot KB / total KB by function")
for f, a in sorted(byfn.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"  {a[1] / 1024:6.1f} / {a[0] / 1024:6.1f}   {a[2] / 1e6:8.0f} M   {f}")
print("== hottest lines by hot bytes")
for (f, c), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"  {a[1]:6d} B hot / {a[0]:6d} B   {a[2] / 1e6:8.0f} M   {f}  {c[0] if c else '?'}:{c[1] if c else 0}")
