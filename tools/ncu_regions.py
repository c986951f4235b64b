"""Attribute ncu warp-stall samples of one kernel to SASS regions / source lines.
  python tools/ncu_regions.py <report.ncu-rep> <object.o> <mangled-kernel-substring>"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
import os

rep, obj, kname = sys.argv[1:4]
tmp = tempfile.mkdtemp()
subprocess.run("cd %s && cuobjdump -xelf all %s >/dev/null 2>&1" % (tmp, os.path.abspath(obj)), shell=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
starts = [i for i, l in enumerate(sass) if ".section\t.text." in l]
st = [i for i in starts if kname in sass[i]][0]
en = min([i for i in starts if i > st] + [len(sass)])
cur, seq = None, []
for l in sass[st:en]:
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), "inlined" in m.group(3))
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((int(m.group(1), 16), cur, m.group(2)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.split("\n")))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) > 10]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = int(data[0][ia], 16)
samp = {int(r[ia], 16) - base: (int(r[isamp] or 0), int(r[iex] or 0)) for r in data}
tot = sum(v[0] for v in samp.values())
print("total samples", tot)
# regions delimited by conv-file source line ranges of the outermost (non-inlined) line info
b, e, lines = collections.Counter(), collections.Counter(), collections.defaultdict(set)
W = 0x200
for o, c, ins in seq:
    s, x = samp.get(o, (0, 0))
    b[o // W] += s
    e[o // W] += x
    if c and c[0].endswith(".cu"):
        lines[o // W].add(c[1])
for k in sorted(b):
    if b[k] >= max(10, tot // 400):
        ls = sorted(lines[k])
        print("%6x samples %6d instr %9d lines %s" % (k * W, b[k], e[k], (ls[:3] + ["..."] + ls[-3:]) if len(ls) > 6 else ls))
print("top instructions:")
for o, c, ins in sorted(seq, key=lambda t: -samp.get(t[0], (0, 0))[0])[:25]:
    print("%6x %6d %9d %-60s %s" % (o, samp.get(o, (0, 0))[0], samp.get(o, (0, 0))[1], ins[:60], c))
