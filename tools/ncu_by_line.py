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
import glob, os
_src = {}
def _line(fn, ln):
    if fn not in _src:
        hits = glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mopa_rl_b200", "csrc", fn))
        _src[fn] = open(hits[0]).read().splitlines() if hits else []
    L = _src[fn]
    return L[ln - 1].strip()[:90] if 0 < ln <= len(L) else ""
order = samp if (len(sys.argv) > 5 and sys.argv[5] == "samples") else inst
for k, _ in order.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 45):
    v = inst[k]
    print("%-24s %6.2f%% %6.2f%% %5.1f  %s" % ("%s:%d" % k, 100 * v / ti, 100 * samp[k] / max(ts, 1), thr[k] / max(v, 1), _line(*k)))
