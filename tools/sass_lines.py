"""Static SASS instruction count per source line of one kernel (nvdisasm -g).  usage: sass_lines.py object.o kernel-substring [min]"""
import os, re, subprocess, sys, tempfile
from collections import defaultdict
obj, kname = sys.argv[1], sys.argv[2]
mn = int(sys.argv[3]) if len(sys.argv) > 3 else 8
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
sec = [t for t in dis.split('//--------------------- .text.')[1:] if t.split(' ')[0].find(kname) >= 0][0]
agg = defaultdict(int); cur = None; tot = 0
for l in sec.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l) and cur: agg[cur] += 1; tot += 1
print("total", tot)
files = defaultdict(int)
for (f, ln), n in sorted(agg.items()):
    files[f] += n
    if n >= mn: print(f"{f:22s}:{ln:4d} {n}")
print(dict(files))
