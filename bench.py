#!/usr/bin/env python
"""Benchmark of the MoPA-RL experience-collection hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload validity|plan|rollout] [--impl reference]

One JSON line on stdout (rank 0).  Under torchrun every rank drives its own GPU; the timed
region is bracketed by barrier + synchronize and the max over ranks is reported.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


# ----------------------------------------------------------------------------- helpers
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return float(json.load(f)["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])), mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def push_setup():
    from helpers import planner_setup
    from mopa_rl_b200.model import load_model

    model = load_model("SawyerPushObstacle-v0")
    ignored, passive, ref = planner_setup(model)
    return model, ignored, passive, ref


def synth_qpos_f32(model, ref, n, seed, pad):
    """BASELINE config 5 queries: active joints ~ U(joint range), passive dims at qpos0; fp32 rows padded to `pad` floats."""
    rng = np.random.Generator(np.random.PCG64(seed))
    jid = [list(model.jnt_qposadr).index(a) for a in ref]
    lo, hi = model.jnt_range[jid, 0].astype(np.float32), model.jnt_range[jid, 1].astype(np.float32)
    q = np.zeros((n, pad), np.float32)
    q[:, :model.nq] = model.qpos0.astype(np.float32)
    q[:, ref] = lo + (hi - lo) * rng.random((n, len(ref)), dtype=np.float32)
    return q


# ----------------------------------------------------------------------------- CPU arms (oracle)
def oracle_validity_rate(model, ignored, q64, threads):
    """Oracle (C restatement of the reference path) on `threads` host threads, one scene per thread."""
    from oracle import oracle

    oracle.build()
    scenes = [oracle.OracleScene(model, ignored, -0.002, "f32") for _ in range(threads)]
    parts = np.array_split(np.arange(len(q64)), threads)
    outs = [None] * threads

    def work(i):
        outs[i] = scenes[i].is_valid(q64[parts[i]])

    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    [t.start() for t in th]
    [t.join() for t in th]
    dt = time.perf_counter() - t0
    return len(q64) / dt, dt, np.concatenate(outs)


def run_reference_arm(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    model, ignored, passive, ref = push_setup()
    cores = os.cpu_count() or 1
    n = args.ref_sample
    rates = []
    for step in range(args.warmup + args.steps):
        q = synth_qpos_f32(model, ref, n, 1234 + step, model.nq)[:, :model.nq].astype(np.float64)
        rate, dt, _ = oracle_validity_rate(model, ignored, q, cores)
        if step >= args.warmup:
            rates.append((rate, dt))
    value = float(np.mean([r for r, _ in rates]))
    line = {
        "impl": "reference", "metric": "state-validity queries/sec (collision-check microbench)", "value": value, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean([d for _, d in rates]) * 1e3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAMES["validity"], "sample_queries_per_step": n},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": "%d queries per step, one oracle scene per host thread (MuJoCo+OMPL cannot be built here: the oracle is the C restatement)" % n},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


WORKLOAD_NAMES = {
    "validity": "config5 collision-check microbench: SawyerPushObstacle-v0, 10M random 7-DoF qpos state-validity queries per GPU, contact_threshold -0.002, cube x {table,bin1} ignored",
}


# ----------------------------------------------------------------------------- GPU arm
def run_validity(args):
    import torch
    import torch.distributed as dist

    from mopa_rl_b200.capi import NativePlanner

    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model, ignored, passive, ref = push_setup()
    planner = NativePlanner(model, passive, ignored, -0.002, 0.1, seed=1234, device=local_rank)
    n = args.queries
    row = ((model.nq + 3) // 4) * 4
    hq = synth_qpos_f32(model, ref, n, 1234 + rank, row)
    h_pinned = torch.from_numpy(hq).pin_memory()
    d_q = h_pinned.to(dev, non_blocking=False)
    d_r = torch.zeros(n, dtype=torch.int32, device=dev)  # uint32 words
    h_words = torch.zeros(n, dtype=torch.int32).pin_memory()
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        planner.is_valid_device(d_q.data_ptr(), row, n, d_r.data_ptr(), 0, stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for k in range(args.steps):
        step()
        ev[k + 1].record()
    barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    # end-to-end: host rows in pinned memory -> result words back in host memory, through the C ABI
    e2e_steps = max(1, min(args.steps, 5))
    planner.is_valid_host_f32(h_pinned.data_ptr(), row, n, h_words.data_ptr(), 0)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        planner.is_valid_host_f32(h_pinned.data_ptr(), row, n, h_words.data_ptr(), 0)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = float(t[0]), float(t[1])
    words = d_r.cpu().numpy().view(np.uint32)
    assert np.array_equal(words & 1, h_words.numpy().view(np.uint32) & 1), "device-resident and end-to-end paths disagree"

    if rank == 0:
        peak, peak_kind = measured_peaks()
        bytes_per_query = row * 4 + 4
        ms_step = total_ms / args.steps
        value = world * n / (ms_step * 1e-3)
        achieved = bytes_per_query * n / (float(np.mean(kernel_ms)) * 1e-3) / 1e9
        # CPU baseline on a bounded sample of the same queries, all host threads, checked against the GPU words
        cores = os.cpu_count() or 1
        ns = min(n, args.cpu_sample)
        rate, dt, ow = oracle_validity_rate(model, ignored, hq[:ns, :model.nq].astype(np.float64), cores)
        mism = int(((ow & 1) != (words[:ns] & 1)).sum())
        rate1, _, _ = oracle_validity_rate(model, ignored, hq[:max(1, ns // cores), :model.nq].astype(np.float64), 1)
        line = {
            "metric": "state-validity queries/sec (collision-check microbench)", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAMES["validity"], "queries_per_gpu": n, "row_bytes": row * 4,
                       "l2": "inputs (%.2f GB per GPU) larger than L2" % (n * row * 4 / 1e9), "valid_fraction": float((words & 1).mean())},
            "clocks": clocks,
            "e2e": {"value": world * n * e2e_steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": n * row * 4, "d2h_bytes_per_step": n * 4,
                    "api": "mopa_is_valid_host_f32 (pinned host rows in, result words out)"},
            "gpu_launches": args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "peak_source": peak_kind, "kernel": "is_valid_kernel", "algorithmic_bytes_per_query": bytes_per_query},
            "cpu_baseline": {"value": rate, "unit": "queries/s", "cores": cores, "kind": "port", "single_thread": rate1,
                             "sample": "%d of the same queries, one oracle scene per host thread" % ns, "gpu_bit_mismatches": mism},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="validity", choices=["validity"])
    ap.add_argument("--queries", type=int, default=10_000_000)
    ap.add_argument("--cpu-sample", type=int, default=2_000_000)
    ap.add_argument("--ref-sample", type=int, default=1_000_000)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    run_validity(args)


if __name__ == "__main__":
    main()
