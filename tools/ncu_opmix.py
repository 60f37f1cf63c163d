"""Summarise an ncu report: SASS opcode mix, top stall lines.  usage: ncu_opmix.py report.ncu-rep [kernel-substr]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
ops, thr = collections.Counter(), collections.Counter()
tot = 0
lines = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[ci["Instructions Executed"]])
    except ValueError:
        continue
    s = r[ci["Source"]].strip().split()
    if not s:
        continue
    op = s[1] if s[0].startswith("@") else s[0]
    ops[op.split(".")[0]] += n
    thr[op.split(".")[0]] += int(r[ci["Thread Instructions Executed"]])
    tot += n
    lines.append((int(r[ci["# Samples"]]), n, r[ci["Source"]].strip(), r[ci["Avg. Threads Executed"]]))
print("total warp-instructions", tot)
for k, v in ops.most_common(28):
    print("%-10s %12d %5.1f%%  avg-threads %.1f" % (k, v, 100 * v / tot, thr[k] / max(v, 1)))
print("\ntop sampled lines")
for smp, n, src, at in sorted(lines, reverse=True)[:25]:
    print(smp, n, at, src)
