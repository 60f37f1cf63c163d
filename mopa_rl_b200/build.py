"""Builds libmopa_b200.so in-tree with nvcc for sm_100a (no torch extension machinery needed:
the boundary is a plain C ABI)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmopa_b200.so")
SOURCES = ["capi.cu", "scene_build.cu", "validity_kernel.cu", "plan.cu", "env_kernel.cu", "env_warp.cu", "rollout.cu", "ik.cu"]
# -fmad=false: the only fused multiply-adds are the explicit fmaf() calls (bit parity with the oracle)
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "--shared", "-Xptxas", "-v"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", f) for f in os.listdir(os.path.join(HERE, "..", "include"))]
    return any(os.path.getmtime(d) > t for d in deps)


# Per-file flags: the collision / planner / rollout-glue kernels must reproduce the oracle bit for bit and are compiled
# without FMA contraction; the fp64 physics kernel is held to 1e-5 on qpos / qvel, not to bit parity, and may contract
# multiply-add pairs (fewer fp64 instructions, shorter dependency chains).
FMAD_ON = {"env_warp.cu"}


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    log, procs = [], []
    for src in SOURCES:
        flags = [f for f in NVCC_FLAGS if f not in ("--shared",)]
        if src in FMAD_ON and os.environ.get("MOPA_ENV_FMAD", "1") != "0":
            flags = [f for f in flags if f != "-fmad=false"]
        if os.environ.get("MOPA_EXTRA_NVCC"):   # diagnostics builds, e.g. MOPA_EXTRA_NVCC=-DMOPA_VK_STATS
            flags += os.environ["MOPA_EXTRA_NVCC"].split()
        if src in FMAD_ON and os.environ.get("MOPA_ENV_MAXREG"):
            flags += ["-maxrregcount=" + os.environ["MOPA_ENV_MAXREG"]]
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + flags + ["-c", "-o", obj, os.path.join(CSRC, src)]
        procs.append((cmd, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for cmd, obj, p in procs:
        out, _ = p.communicate()
        log.append(" ".join(cmd) + "\n" + out)
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed building " + obj)
        objs.append(obj)
    cmd = [nvcc, "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append(" ".join(cmd) + "\n" + r.stdout)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout)
    if r.returncode:
        raise RuntimeError("nvcc failed linking libmopa_b200.so")
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write("\n".join(log))
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)
