"""Top SASS instructions by a stall reason, with the CUDA line they map to.
usage: ncu_stalls.py report.ncu-rep cubin kernel_substr stall_column [top]"""
import collections, csv, io, re, subprocess, sys
rep, cubin, kname, col = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], stdout=subprocess.PIPE, text=True).stdout
omap, cur, infn = {}, ("?", 0), False
for ln in dis.splitlines():
    if ln.startswith(".text.") or re.match(r"^\s*\.section\s+\.text\.", ln):
        infn = kname in ln; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m: omap[int(m.group(1), 16)] = cur
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
sass = [r for r in rows[2:] if len(r) >= len(hdr) and r[0].startswith("0x")]
base = int(sass[0][0], 16)
tot = sum(int(r[ci[col]] or 0) for r in sass)
byline = collections.Counter()
for r in sass:
    byline[omap.get(int(r[0], 16) - base, ("?", 0))] += int(r[ci[col]] or 0)
print("total", col, tot)
for r in sorted(sass, key=lambda r: -int(r[ci[col]] or 0))[:top]:
    k = omap.get(int(r[0], 16) - base, ("?", 0))
    print("%6d %5.2f%%  %-22s %s" % (int(r[ci[col]]), 100 * int(r[ci[col]]) / max(tot, 1), "%s:%d" % k, r[1][:70]))
print("--- by line")
for k, v in byline.most_common(top):
    print("%6d %5.2f%%  %s:%d" % (v, 100 * v / max(tot, 1), k[0], k[1]))
