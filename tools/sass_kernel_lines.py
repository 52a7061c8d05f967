"""SASS bytes per source-line bucket for ONE kernel's .text (incl. its device functions).  usage: obj kernel-substr file bucket"""
import re, collections, subprocess, sys, tempfile, os, glob
obj = os.path.abspath(sys.argv[1]); want = sys.argv[2]; f0 = sys.argv[3]; bk = int(sys.argv[4]) if len(sys.argv) > 4 else 20
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, capture_output=True)
cub = glob.glob(os.path.join(d, "*.cubin"))[0]
out = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
cur = None; cnt = collections.Counter(); on = False; infunc = False
for ln in out.splitlines():
    m = re.match(r'\.text\.(\S+):', ln)
    if m: on = want in m.group(1); infunc = False; continue
    if re.match(r'\$\S+:', ln): infunc = True
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if on and not infunc and re.match(r'\s+/\*[0-9a-f]{4,6}\*/', ln) and cur: cnt[cur] += 16
byfile = collections.Counter()
for (f, l), c in cnt.items(): byfile[f] += c
print("main body by file:", byfile.most_common())
b = collections.Counter()
for (f, l), c in cnt.items():
    if f == f0: b[l // bk * bk] += c
print(" ".join(f"{k}:{b[k]}" for k in sorted(b)))
