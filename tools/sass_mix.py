#!/usr/bin/env python3
"""Static instruction mix per (outermost) source line of a kernel, from an object file compiled with -lineinfo.
usage: sass_mix.py <object.o> <mangled kernel> <file suffix> [line ...]"""
import collections, os, re, subprocess, sys, tempfile
obj, kern, suffix = sys.argv[1:4]
want = [int(x) for x in sys.argv[4:]]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, stdout=subprocess.DEVNULL)
cubin = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], stdout=subprocess.PIPE, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith(".text." + kern + ":"))
cur = None
hist = collections.defaultdict(collections.Counter)
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        chain = [(m.group(1), int(m.group(2)))] + [(a, int(b)) for a, b in re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))]
        outer = [c for c in chain if c[0].endswith(suffix)]
        cur = outer[-1][1] if outer else None
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)', l)
    if m:
        hist[cur][m.group(3)] += 1
tot = sum(sum(c.values()) for c in hist.values())
print("total static instructions", tot)
for ln in sorted(hist, key=lambda x: (x is None, x)):
    c = hist[ln]
    n = sum(c.values())
    if (want and ln in want) or (not want and n * 100 >= tot):
        print("line %5s  %5d  %s" % (ln, n, " ".join("%s:%d" % kv for kv in c.most_common(10))))
