"""Instruction / stall-sample share per phase of is_valid_kernel (validity_kernel.cu): every SASS instruction is attributed to the
outermost validity_kernel.cu line of its inlined-at chain.  usage: ncu_vk_phases.py report.ncu-rep validity_kernel.cubin [kernel_substr]"""
import collections, csv, io, re, subprocess, sys

rep, cubin = sys.argv[1:3]
kname = sys.argv[3] if len(sys.argv) > 3 else "is_valid_kernelILi128ELb0"
INNER = len(sys.argv) > 5 and sys.argv[5] == "inner"   # attribute to the innermost line of the kernel's file (looks inside the lambdas)
src = open(__file__.rsplit("/", 2)[0] + "/mopa_rl_b200/csrc/validity_kernel.cu").read().splitlines()
def find(tag):
    return next(i + 1 for i, l in enumerate(src) if tag in l)
marks = [(find("auto run_queue"), "D portal refinement (queue)"), (find("for (int tile = blockIdx.x"), "tile prologue"),
         (find("fk_state(S, rq"), "A forward kinematics"), (find("const uint32_t live = q < n"), "B sweep: flush / push"),
         (find("for (int g = 0; g < ngroup"), "B sweep: cull tests"), (find("---- phase C: analytic"), "C analytic + box-box"),
         (find("portal-refinement candidates: conservative"), "C portal pre-test + queue"), (find("out[q] = res[tid]"), "tile epilogue / result words")]
marks.sort()
def phase(line):
    name = "prologue"
    for l0, nm in marks:
        if line >= l0:
            name = nm
    return name
dis = subprocess.run(["nvdisasm", "-gi", cubin], stdout=subprocess.PIPE, text=True).stdout
omap, cur, infn, block = {}, None, False, []
for ln in dis.splitlines():
    if re.match(r"^\s*\.section\s", ln) or ln.startswith(".text."):
        infn = ".text." in ln and kname in ln
        continue
    if not infn:
        continue
    if "//## File" in ln:
        block += re.findall(r'"([^"]+)", line (\d+)', ln)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        if block:   # the inlined-at chain of this instruction: innermost first; keep the outermost line of the kernel's own file
            outer = [int(l) for f, l in block if f.endswith("validity_kernel.cu")]
            cur = (outer[0] if INNER else outer[-1]) if outer else cur
            block = []
        omap[int(m.group(1), 16)] = cur
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
sass = [r for r in rows[2:] if len(r) >= len(hdr) and r[0].startswith("0x")]
base = int(sass[0][0], 16)
ic = next(i for h, i in ci.items() if h.startswith("# Instructions Executed") or h == "Instructions Executed")
tc = next(i for h, i in ci.items() if "Thread Instructions Executed" in h)
sc = next(i for h, i in ci.items() if h.startswith("# Samples") or h == "Samples" or "Sampling Data (All)" in h)
inst, samp, thr = collections.Counter(), collections.Counter(), collections.Counter()
linst, lsamp, lthr = collections.Counter(), collections.Counter(), collections.Counter()
for r in sass:
    line = omap.get(int(r[0], 16) - base)
    ph = phase(line) if line else "?"
    inst[ph] += int(r[ic] or 0); thr[ph] += int(r[tc] or 0); samp[ph] += int(r[sc] or 0)
    linst[line] += int(r[ic] or 0); lthr[line] += int(r[tc] or 0); lsamp[line] += int(r[sc] or 0)
ti, ts = sum(inst.values()), sum(samp.values())
print("total warp-inst %d, samples %d" % (ti, ts))
print("%-34s %7s %7s %9s" % ("phase", "inst%", "samp%", "thr/inst"))
for ph, _ in sorted(inst.items(), key=lambda kv: -samp[kv[0]]):
    print("%-34s %6.1f%% %6.1f%% %9.1f" % (ph, 100 * inst[ph] / ti, 100 * samp[ph] / ts, thr[ph] / max(inst[ph], 1)))

if len(sys.argv) > 4:   # per outermost source line
    print()
    for line, _ in sorted(linst.items(), key=lambda kv: -kv[1])[:int(sys.argv[4])]:
        print("%5s %6.2f%% %6.2f%% %5.1f  %s" % (line, 100 * linst[line] / ti, 100 * lsamp[line] / ts, lthr[line] / max(linst[line], 1), src[line - 1].strip()[:110] if line else ""))
