"""Code size of every entry / device function in an object file (bytes), from `nvdisasm`-style labels in cuobjdump -sass."""
import re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
cur, sizes, last = None, {}, 0
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1); sizes[cur] = 0; continue
    m = re.match(r"\s*(\$\S+|\.L_x_\d+):", ln)
    m2 = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(\S.*?);", ln)
    if m2 and cur:
        sizes[cur] = int(m2.group(1), 16) + 16
for k, v in sizes.items():
    print(f"{v:8d}  {k[:100]}")
