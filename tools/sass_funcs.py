"""Sizes of the device functions laid out inside each kernel's .text (nvdisasm labels).  usage: sass_funcs.py obj.o [kernel-substring]"""
import re, subprocess, sys, tempfile, os, glob
obj = os.path.abspath(sys.argv[1]); want = sys.argv[2] if len(sys.argv) > 2 else ""
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, capture_output=True)
cub = glob.glob(os.path.join(d, "*.cubin"))[0]
out = subprocess.run(["nvdisasm", "-c", cub], capture_output=True, text=True).stdout
kern = None; marks = []; last = 0
for ln in out.splitlines():
    m = re.match(r'\.text\.(\S+):', ln)
    if m:
        if kern and marks: 
            marks.append(("<end>", last + 16)); 
            if want in kern:
                print("==", kern[:90])
                for (a, s), (b, e) in zip(marks, marks[1:]): print(f"  {e - s:7d}  {a[:110]}")
        kern = m.group(1); marks = [("<main body>", 0)]; last = 0; continue
    m = re.match(r'(\$\S+):', ln)
    if m and kern:
        marks.append((m.group(1).split('$')[-1], last + 16)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/', ln)
    if m: last = int(m.group(1), 16)
if kern and marks and want in kern:
    marks.append(("<end>", last + 16))
    print("==", kern[:90])
    for (a, s), (b, e) in zip(marks, marks[1:]): print(f"  {e - s:7d}  {a[:110]}")
