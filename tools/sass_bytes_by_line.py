"""Static code size per source line of one kernel (nvdisasm -g line markers): where the instruction bytes are.
    python tools/sass_bytes_by_line.py obj.o kernel-substring [top_n]"""
import collections, glob, os, re, subprocess, sys, tempfile
obj, want = os.path.abspath(sys.argv[1]), sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, capture_output=True)
out = subprocess.run(["nvdisasm", "-g", "-c", glob.glob(os.path.join(d, "*.cubin"))[0]], capture_output=True, text=True).stdout
kern, cur, cnt, fn = None, None, collections.Counter(), "<main body>"
byfn = collections.Counter()
for ln in out.splitlines():
    m = re.match(r'\.text\.(\S+):', ln)
    if m: kern = m.group(1); fn = "<main body>"; continue
    if not kern or want not in kern: continue
    m = re.match(r'(\$\S+):', ln)
    if m: fn = m.group(1).split('$')[-1][-40:]; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,6}\*/', ln):
        cnt[(fn, cur)] += 16; byfn[fn] += 16
print("== bytes by function"); 
for f, b in byfn.most_common(12): print(f"{b:7d}  {f}")
print("== top lines of <main body>")
for (f, c), b in [x for x in cnt.most_common() if x[0][0] == "<main body>"][:top]:
    print(f"{b:6d}  {c[0]}:{c[1]}" if c else f"{b:6d}  ?")
