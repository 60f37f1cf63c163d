"""Dynamic instruction / stall-sample share per stage of env_warp.cu's w_substep.
usage: ncu_by_stage.py report.ncu-rep env_warp.cubin"""
import collections, csv, io, re, subprocess, sys

rep, cubin = sys.argv[1:3]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], stdout=subprocess.PIPE, text=True).stdout
omap, cur, inl, infn = {}, ("?", 0), "", False
for ln in dis.splitlines():
    if re.match(r"^\s*\.section\s+\.text\.", ln) or ln.startswith(".text."):
        infn = "env_step_warp" in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        # inlined-at chain: take the outermost env_warp.cu line if present
        chain = re.findall(r'File "([^"]+)", line (\d+)', ln)
        outer = [(f.split("/")[-1], int(l)) for f, l in chain if f.endswith("env_warp.cu")]
        inl = outer[-1] if outer else None
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        omap[int(m.group(1), 16)] = (cur, inl)
STAGES = [(83, 102, "w_chol"), (103, 121, "w_solve"), (122, 124, "pair_contacts"), (125, 134, "kbi"), (144, 214, "kinematics"),
          (215, 235, "inertia"), (236, 246, "velocity chain"), (247, 294, "RNE"), (295, 320, "CRBA"), (321, 343, "forces+qacc0"),
          (344, 369, "limit rows"), (370, 432, "broadphase"), (433, 487, "narrowphase glue"), (488, 540, "row jacobians"),
          (541, 573, "Y / A"), (574, 633, "PGS"), (634, 652, "J^T f"), (653, 682, "integrate"), (683, 900, "env epilogue/prologue")]
def stage(key):
    (f, l), outer = key
    if f != "env_warp.cu":
        if f == "contact.cuh": return "contact routines"
        if outer: f, l = outer
        else: return f
    for a, b, name in STAGES:
        if a <= l <= b: return name
    return "other env_warp"
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
sass = [r for r in rows[2:] if len(r) >= len(hdr) and r[0].startswith("0x")]
base = int(sass[0][0], 16)
inst, samp, thr = collections.Counter(), collections.Counter(), collections.Counter()
for r in sass:
    st = stage(omap.get(int(r[0], 16) - base, (("?", 0), None)))
    n = int(r[ci["Instructions Executed"]])
    inst[st] += n
    thr[st] += int(r[ci["Thread Instructions Executed"]])
    samp[st] += int(r[ci["# Samples"]])
ti, ts = sum(inst.values()), sum(samp.values())
print("total warp-inst %d  samples %d" % (ti, ts))
for k, v in inst.most_common():
    print("%-24s inst %6.2f%%  samples %6.2f%%  thr/inst %5.1f" % (k, 100 * v / ti, 100 * samp[k] / ts, thr[k] / max(v, 1)))
