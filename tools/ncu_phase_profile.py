#!/usr/bin/env python3
"""Attributes the warp-state samples of an ncu source-page export (SASS view, CSV) to source lines of the
OUTERMOST frame in a given file, using nvdisasm -gi line info of the same cubin.
usage: ncu_phase_profile.py <sass.csv> <nvdisasm -gi output> <mangled kernel name> <file suffix>"""
import collections
import csv
import re
import sys

sass_csv, dis, kern, fsuffix = sys.argv[1:5]
# --- line info per instruction offset
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kern + ":"))
off2line = {}
cur_outer, cur_inner = None, None
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        f, ln, rest = m.group(1), int(m.group(2)), m.group(3)
        cur_inner = (f, ln)
        chain = [(f, ln)] + [(a, int(b)) for a, b in re.findall(r'inlined at "([^"]+)", line (\d+)', rest)]
        cur_chain = chain
        outer = [c for c in chain if c[0].endswith(fsuffix)]
        cur_outer = outer[-1][1] if outer else None
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        off2line[int(m.group(1), 16)] = (cur_outer, cur_inner, m.group(2).strip())
rows = list(csv.reader(open(sass_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ci = {n: i for i, n in enumerate(hdr)}
data = rows[h + 1:]
base = int(data[0][0], 16) if data[0][0].startswith("0x") else int(data[0][0])
per_line = collections.Counter()
per_line_inst = collections.Counter()
stall_by_line = collections.defaultdict(collections.Counter)
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = 0
for r in data:
    if not r or not r[0]:
        continue
    a = int(r[0], 16) if r[0].startswith("0x") else int(r[0])
    info = off2line.get(a - base)
    ln = info[0] if info else None
    s = int(r[ci["# Samples"]] or 0)
    per_line[ln] += s
    per_line_inst[ln] += int(r[ci["Instructions Executed"]] or 0)
    for n in stalls:
        v = int(r[ci[n]] or 0)
        if v:
            stall_by_line[ln][n] += v
    tot += s
print("total samples", tot, "total warp instructions", sum(per_line_inst.values()))
for ln, s in sorted(per_line.items(), key=lambda t: (t[0] is None, t[0])):
    if s * 200 < tot and per_line_inst[ln] * 200 < sum(per_line_inst.values()):
        continue
    top = ", ".join("%s %.0f%%" % (k.replace("stall_", ""), 100 * v / max(s, 1)) for k, v in stall_by_line[ln].most_common(3))
    print("line %5s  samples %6.2f%%  instr %6.2f%%   %s" % (ln, 100 * s / tot, 100 * per_line_inst[ln] / sum(per_line_inst.values()), top))
