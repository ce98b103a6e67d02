#!/usr/bin/env python
"""Attribute ncu SASS-level samples to source lines of INLINED headers (ncu's CUDA source page only
carries the .cu file).  usage: sass_lines.py prof.ncu-rep lib.so 'StageKernelILi0ELi4ELi3' [launch]"""
import collections
import glob
import os
import csv
import re
import subprocess
import sys
import tempfile

rep, lib, pat = sys.argv[1:4]
which = sys.argv[4] if len(sys.argv) > 4 else "0"
KIND = sys.argv[5] if len(sys.argv) > 5 else "phase_kernel"
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
import glob
import os
cub = max(glob.glob(os.path.join(d, "*.cubin")), key=os.path.getsize)
dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
line_of = {}
cur_fn, cur_loc, active = None, None, False
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+)", ln) or re.match(r"\s*//-+ \.text\.(\S+)", ln)
    if m:
        active = pat in m.group(1) and KIND in m.group(1)
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_loc = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur_loc, m.group(2))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", which, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]
iA, iS, iN = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed")
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
base = None
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
ops = collections.Counter()
for r in rows[2:]:
    if len(r) < len(h) or not r[iA].startswith('0x'):
        continue
    a = int(r[iA], 16)
    if base is None:
        base = a
    loc, txt = line_of.get(a - base, (("?", 0), "?"))
    key = loc or ("?", 0)
    s, n = int(r[iS]), int(r[iN])
    agg[key]["samples"] += s
    agg[key]["inst"] += n
    tot["samples"] += s
    tot["inst"] += n
    toks = [x for x in r[h.index("Source")].split() if not x.startswith("@")]
    ops[toks[0] if toks else "?"] += n
    for c in stalls:
        v = int(r[h.index(c)] or 0)
        agg[key][c] += v
    for c in ("L1 Wavefronts Shared Excessive", "L1 Wavefronts Shared"):
        agg[key][c] += int(r[h.index(c)] or 0)
        tot[c] += int(r[h.index(c)] or 0)
print("total samples %d, warp-level instructions %d" % (tot["samples"], tot["inst"]))
print("%-22s %7s %7s  top stalls" % ("file:line", "smpl%", "inst%"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:40]:
    top = sorted(((c, v[c]) for c in stalls), key=lambda x: -x[1])[:3]
    print("%-22s %6.1f%% %6.1f%%  %s" % ("%s:%d" % k, 100.0 * v["samples"] / tot["samples"], 100.0 * v["inst"] / tot["inst"],
                                       ", ".join("%s=%d" % (c[6:], n) for c, n in top if n)))
print("-- shared-memory wavefronts: total %d, excessive %d" % (tot["L1 Wavefronts Shared"], tot["L1 Wavefronts Shared Excessive"]))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["L1 Wavefronts Shared Excessive"])[:10]:
    print("   %-22s wavefronts %9d excessive %9d" % ("%s:%d" % k, v["L1 Wavefronts Shared"], v["L1 Wavefronts Shared Excessive"]))
print("-- opcode mix (warp instr %)")
for o, n in ops.most_common(22):
    print("   %-14s %5.1f%%" % (o, 100.0 * n / tot["inst"]))
