"""Dynamic warp instructions per 128-position group, by source line, from an ncu report + nvdisasm -g -c listing.
usage: python tools/ncu_groups.py report.ncu-rep object.o kernel-name-substring n_groups [min]"""
import csv, re, subprocess, sys, os, tempfile
from collections import defaultdict
rep, obj, kname, groups = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
mn = float(sys.argv[5]) if len(sys.argv) > 5 else 2.0
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
sec = [t for t in dis.split('//--------------------- .text.')[1:] if t.split(' ')[0].find(kname) >= 0][0]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ia, ii = hdr.index("Address"), hdr.index("Instructions Executed")
data = rows[2:]; base = int(data[0][ia], 16)
line_of = {}; cur = None
for l in sec.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m and cur: line_of[int(m.group(1), 16)] = cur
agg = defaultdict(lambda: [0, 0]); tot = 0
for r in data:
    off = int(r[ia], 16) - base; key = line_of.get(off, ("?", 0)); n = int(r[ii] or 0)
    agg[key][0] += n; agg[key][1] += 1; tot += n
print("per group total %.1f" % (tot / groups))
for (f, ln), (n, c) in sorted(agg.items()):
    if n / groups >= mn: print(f"{f:20s}:{ln:4d} dyn/group {n/groups:7.1f}  static {c}")
