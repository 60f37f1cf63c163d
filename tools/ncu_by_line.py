"""Attribute executed warp-instructions and stall samples of an ncu report to CUDA source lines
(via nvdisasm --print-line-info on the cubin).  usage: ncu_by_line.py report.ncu-rep cubin kernel_substr"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, cubin, kname = sys.argv[1:4]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], stdout=subprocess.PIPE, text=True).stdout
# walk the function: track current //## File "...", line N   and instruction offsets /*0010*/
in_fn = False
cur = ("?", 0)
inline_stack = ""
offs = []  # (offset, file:line)
for ln in dis.splitlines():
    if ln.startswith(".text.") or re.match(r"^\s*\.section\s+\.text\.", ln):
        in_fn = kname in ln
        continue
    if not in_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        inline_stack = m.group(3)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
    if m:
        offs.append((int(m.group(1), 16), cur))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
sass = [r for r in rows[2:] if len(r) >= len(hdr) and r[0].startswith("0x")]
base = int(sass[0][0], 16)
omap = dict(offs)
inst, samp, thr = collections.Counter(), collections.Counter(), collections.Counter()
for r in sass:
    off = int(r[0], 16) - base
    key = omap.get(off, ("?", 0))
    n = int(r[ci["Instructions Executed"]])
    inst[key] += n
    thr[key] += int(r[ci["Thread Instructions Executed"]])
    samp[key] += int(r[ci["# Samples"]])
ti, ts = sum(inst.values()), sum(samp.values())
print("total warp-inst %d, samples %d, mapped offsets %d/%d" % (ti, ts, sum(1 for r in sass if (int(r[0], 16) - base) in omap), len(sass)))
byfile = collections.Counter()
for k, v in inst.items():
    byfile[k[0]] += v
print("by file:", {k: "%.1f%%" % (100 * v / ti) for k, v in byfile.most_common()})
print("%-28s %8s %8s %8s" % ("file:line", "inst%", "samp%", "thr/inst"))
for k, v in inst.most_common(45):
    print("%-28s %7.2f%% %7.2f%% %8.1f" % ("%s:%d" % k, 100 * v / ti, 100 * samp[k] / max(ts, 1), thr[k] / max(v, 1)))
