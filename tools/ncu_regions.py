"""Executed warp-instructions and stall samples of one kernel in an ncu report, grouped by source function.
usage: python tools/ncu_regions.py REPORT.ncu-rep [top_lines]"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None
lines = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]
        continue
    if r[0].isdigit():
        try:
            s = int(r[4]); inst = int(r[7])
        except ValueError:
            continue
        lines.append((inst, s, cur, int(r[0]), r[1].strip()))
# map line -> enclosing function by scanning the source files
funcs = {}
for f in set(l[2] for l in lines):
    try:
        src = open(f).read().splitlines()
    except OSError:
        continue
    name, table = "?", []
    for i, t in enumerate(src, 1):
        m = re.match(r"^(?:template.*>\s*)?(?:__global__|__device__|static|inline).*?\b(ehb_\w+)\s*\(", t)
        if m:
            name = m.group(1)
        table.append(name)
    funcs[f] = table
agg = collections.Counter(); aggs = collections.Counter()
for inst, s, f, ln, src in lines:
    t = funcs.get(f)
    k = t[ln - 1] if t and ln - 1 < len(t) else f.split("/")[-1]
    agg[k] += inst; aggs[k] += s
tot = sum(agg.values()) or 1; tots = sum(aggs.values()) or 1
print("total warp-instructions %d, samples %d" % (tot, tots))
for k, v in agg.most_common():
    print("%-28s inst %9d %5.1f%%   samples %5d %5.1f%%" % (k, v, 100 * v / tot, aggs[k], 100 * aggs[k] / tots))
lines.sort(reverse=True)
for inst, s, f, ln, src in lines[:top]:
    print("%8d %5d %s:%d %s" % (inst, s, f.split("/")[-1], ln, src[:120]))
