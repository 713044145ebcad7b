"""Attribute ncu's per-SASS-instruction counters to source lines.
usage: python tools/ncu_lines.py report.ncu-rep disasm.txt(nvdisasm -g -c of the same cubin) [file-substring]
Prints, per source line, executed warp instructions, stall samples and shared-memory wavefronts."""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, dis = sys.argv[1], sys.argv[2]
only = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
iw = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
iwi = hdr.index("L1 Wavefronts Shared Ideal") if "L1 Wavefronts Shared Ideal" in hdr else None
ith = hdr.index("Avg. Threads Executed")
data = rows[2:]
base = int(data[0][ia], 16)
# nvdisasm: lines like "        /*0010*/  INSTR ;" preceded by //## File "...", line N (inlined chains also appear)
line_of = {}
cur = None
for l in open(dis):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
agg = defaultdict(lambda: [0, 0, 0, 0])
tot_i = tot_s = 0
for r in data:
    off = int(r[ia], 16) - base
    key = line_of.get(off, ("?", 0))
    n = int(r[ii] or 0)
    s = int(r[isamp] or 0)
    agg[key][0] += n
    agg[key][1] += s
    if iw is not None:
        agg[key][2] += int(r[iw] or 0)
        agg[key][3] += int(r[iwi] or 0)
    tot_i += n
    tot_s += s
print(f"total warp instructions {tot_i}, samples {tot_s}")
src_cache = {}
for (f, ln), (n, s, w, wi) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    if only and only not in f:
        continue
    print(f"{f:18s}:{ln:4d}  inst {100.0 * n / tot_i:5.1f}%  samples {100.0 * s / max(tot_s, 1):5.1f}%  smem wf {w:>10d} (ideal {wi})")
