#!/usr/bin/env python
"""Benchmark of the MoPA-RL experience-collection hot path on B200 ("Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload rollout|validity] [--impl reference]

rollout  (default, BASELINE.json metric): env-steps/sec incl. planner, SawyerPushObstacle-v0 MoPA,
         4096 vectorised envs per GPU.  One "step" = one tick of the vectorised runner = one env.step
         (75 physics substeps) for every env, plus the policy / validity / RRT work of the envs that
         finished their macro action.  With N > 1 every rank owns its own env shard and the ranks
         all-gather the new transition records each tick (NCCL) into a replicated replay.
validity BASELINE config 5: 10M random Sawyer qpos state-validity queries per GPU.
One JSON line on stdout (rank 0); max over ranks of device-timed regions bracketed by barriers.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "env-steps/sec (incl. planner) SawyerPushObstacle-v0"
METRIC_ASSEMBLY = "env-steps/sec (incl. planner) SawyerAssemblyObstacle-v0"
TASK_ENV = {"push": "SawyerPushObstacle-v0", "assembly": "SawyerAssemblyObstacle-v0", "lift": "SawyerLiftObstacle-v0", "lift-ik": "SawyerLiftObstacle-v0",
            "pusher": "PusherObstacle-v0"}


# scripts/3d/{push,assembly,lift}/mopa.sh: omega per task (action_range 0.5, reuse_data, max_reuse_data 15 in all three)
TASK_OMEGA = {"push": 0.7, "assembly": 0.7, "lift": 0.5, "lift-ik": 0.05, "pusher": 0.5}


def task_config(task, max_iter):
    from mopa_rl_b200.rollout import MoPAConfig

    if task == "pusher":
        # BASELINE configs[0]: scripts/2d/mopa.sh (omega 0.5, action_range 1.0, reuse_data, max_reuse_data 30) + config/pusher.py (range 0.2,
        # simple_planner_range 0.1, contact_threshold -0.0015, step_size 0.04, joint_margin 0; timelimit 1.0 s / simple 0.02 s -> half the
        # iteration cap of the 2.0 s Sawyer presets, and 1 % of it for the simple planner)
        return MoPAConfig(omega=0.5, action_range=1.0, ac_scale=0.1, step_size=0.04, joint_margin=0.0, contact_threshold=-0.0015, range=0.2,
                          simple_planner_range=0.1, max_iter=max(1, max_iter // 2), simple_max_iter=max(1, max_iter // 100), reuse_data=True,
                          max_reuse_data=30)
    if task == "lift-ik":
        # BASELINE configs[2]: scripts/3d/lift/mopa_ik.sh (MoPA-SAC IK: use_ik_target, ik_target grip_site, action_range 0.2, omega 0.05)
        return MoPAConfig(max_iter=max_iter, reuse_data=True, max_reuse_data=15, omega=0.05, action_range=0.2, use_ik_target=True, ik_target="grip_site")
    return MoPAConfig(max_iter=max_iter, reuse_data=True, max_reuse_data=15, omega=TASK_OMEGA[task])


def task_env_class(task):
    from mopa_rl_b200 import envs

    return {"assembly": envs.VecSawyerAssemblyObstacle, "lift": envs.VecSawyerLiftObstacle, "lift-ik": envs.VecSawyerLiftObstacle,
            "pusher": envs.VecPusherObstacle}.get(task, envs.VecSawyerPushObstacle)


TASK_ADIM = {"push": 7, "assembly": 7, "lift": 8, "lift-ik": 8, "pusher": 4}


def task_metric(task):
    return METRIC.replace("SawyerPushObstacle-v0", TASK_ENV[task])
WORKLOADS = {
    "rollout": "SawyerPushObstacle-v0 MoPA (omega OMEGA, action_range 0.5, RRT-Connect range 0.1), %d vectorised envs per GPU, uniform random-exploration policy",
    "validity": "config5 collision-check microbench: SawyerPushObstacle-v0, %d random 7-DoF qpos state-validity queries per GPU, contact_threshold -0.002, cube x {table,bin1} ignored",
}


def ncu_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture (or None)."""
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return int(json.load(f)["dram_bytes_per_launch"])
    except Exception:
        return None


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
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
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


def push_setup(task="push"):
    from mopa_rl_b200.model import load_model
    from mopa_rl_b200.rollout import planner_inputs

    model = load_model(TASK_ENV[task])
    ignored, passive, ref = planner_inputs(model)
    return model, ignored, passive, ref


def synth_qpos_f32(model, ref, n, seed, pad):
    rng = np.random.Generator(np.random.PCG64(seed))
    jid = [list(model.jnt_qposadr).index(a) for a in ref]
    lo, hi = model.jnt_range[jid, 0].astype(np.float32), model.jnt_range[jid, 1].astype(np.float32)
    q = np.zeros((n, pad), np.float32)
    q[:, :model.nq] = model.qpos0.astype(np.float32)
    q[:, ref] = lo + (hi - lo) * rng.random((n, len(ref)), dtype=np.float32)
    return q


# ----------------------------------------------------------------------------- CPU arms (the oracle = C/numpy restatement
# of the reference path; MuJoCo 2.0 + OMPL cannot be built here)
def _cpu_rollout_worker(args):
    gid, macros, seed, max_iter, task = args
    sys.path.insert(0, ROOT)
    from mopa_rl_b200 import rng
    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.model import load_model
    from mopa_rl_b200.rollout import env_planner_inputs
    from oracle.rollout_oracle import ScalarMoPARunner

    cls = task_env_class(task)
    model = load_model(cls.ENV_ID)
    ignored, passive, _ = env_planner_inputs(cls, model)
    adim = TASK_ADIM[task]

    def policy(g, k):
        u = rng.uniform01(seed + 7, np.uint64(g), np.uint64(k), np.arange(adim, dtype=np.uint64))
        return (2.0 * u - 1.0).astype(np.float32)

    r = ScalarMoPARunner(model, DynModel(model), task_config(task, max_iter), ignored, passive, gid, seed, policy, task="lift" if task == "lift-ik" else task,
                         max_episode_steps=400 if task == "pusher" else 250)
    t0 = time.perf_counter()
    for _ in range(macros):
        r.macro_step()
    return r.env_steps, time.perf_counter() - t0


def cpu_rollout_rate(cores, macros, seed, max_iter, base_gid=0, task="push"):
    """One scalar runner (env + planner, like one MPI rank of the reference) per host core."""
    from oracle import oracle

    oracle.build()
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_rollout_worker, [(base_gid + i, macros, seed, max_iter, task) for i in range(cores)])
    wall = time.perf_counter() - t0
    steps = sum(s for s, _ in res)
    busy = max(t for _, t in res)
    return steps / busy, steps, busy, wall


def oracle_validity_rate(model, ignored, q64, threads):
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
    cores = os.cpu_count() or 1
    if args.workload == "validity":
        model, ignored, passive, ref = push_setup()
        rates = []
        for step in range(args.warmup + args.steps):
            q = synth_qpos_f32(model, ref, args.ref_sample, 1234 + step, model.nq)[:, :model.nq].astype(np.float64)
            rate, dt, _ = oracle_validity_rate(model, ignored, q, cores)
            if step >= args.warmup:
                rates.append((rate, dt))
        value, ms = float(np.mean([r for r, _ in rates])), float(np.mean([d for _, d in rates]) * 1e3)
        metric, unit, sample = "state-validity queries/sec (collision-check microbench)", "queries/s", "%d queries per step" % args.ref_sample
        cfg = {"workload": WORKLOADS["validity"] % args.queries}
    else:
        rates = []
        for step in range(args.warmup + args.steps):
            rate, steps, busy, wall = cpu_rollout_rate(cores, args.cpu_macros, 1234, args.max_iter, base_gid=1000 * step, task=args.task)
            if step >= args.warmup:
                rates.append((rate, busy))
        value, ms = float(np.mean([r for r, _ in rates])), float(np.mean([d for _, d in rates]) * 1e3)
        metric, unit = task_metric(args.task), "env-steps/s"
        sample = "%d macro actions per scalar runner per step, one runner (env + planner) per host core" % args.cpu_macros
        cfg = rollout_config(args, args.envs)
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port",
                         "sample": sample + " (oracle = C/numpy restatement of the reference path; MuJoCo 2.0 + OMPL cannot be built here)"},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


# ----------------------------------------------------------------------------- GPU arm: rollout
class HostLoopPolicy:
    """End-to-end arm: the policy lives on the host (as the reference's actor loop does): every tick the
    observations of all environments come back to pinned host memory and the actions go up from pinned memory."""

    def __init__(self, torch, device, seed, n, adim=7):
        self.torch, self.device, self.n, self.adim = torch, device, n, adim
        self.rng = np.random.Generator(np.random.PCG64(seed))
        self.h_obs = torch.zeros(n, 40, dtype=torch.float32).pin_memory()
        self.h_act = torch.zeros(n, adim, dtype=torch.float32).pin_memory()
        self.h2d = self.d2h = 0

    def __call__(self, obs, env_ids=None, macro_index=None):
        self.h_obs.copy_(obs, non_blocking=False)
        self.h_act.copy_(self.torch.from_numpy(self.rng.uniform(-1, 1, (self.n, self.adim)).astype(np.float32)))
        self.d2h += self.n * 40 * 4
        self.h2d += self.n * self.adim * 4
        return self.h_act.to(self.device, non_blocking=True)


def run_rollout(args):
    import torch
    import torch.distributed as dist

    from mopa_rl_b200.replay import ReplicatedReplay
    from mopa_rl_b200.rollout import NativeMoPARolloutRunner

    env_cls = task_env_class(args.task)

    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.envs
    cfg = task_config(args.task, args.max_iter)   # scripts/3d/<task>/mopa.sh
    slab = args.slab or max(1024, n)   # steady state emits ~0.67 n records per tick (main + relabelled, push preset); bursts queue up and drain

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make(policy=None):
        venv = env_cls(n, seed=1234, device=local_rank, env_id_offset=rank * n)
        return NativeMoPARolloutRunner(venv, cfg, policy=policy)

    def timed(runner, replay, h_block=None):
        """`settle` untimed ticks (start-up transient: every env plans at tick 0), W warm-up ticks, then K timed ticks.
        Returns (device ms, wall ms, env_steps, launches, env-kernel ms, d2h bytes)."""
        d2h = 0
        for _ in range(args.settle + args.warmup):
            runner.tick()
            replay.exchange(runner)
        barrier()
        l0, s0 = runner.launches, runner.env_steps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            runner.tick()
            replay.exchange(runner)
            if h_block is not None:   # end-to-end arm: this tick's transition records (header row + records) back to pinned host memory
                h_block.copy_(replay.last_block, non_blocking=True)
                d2h += h_block.numel() * 4
        replay.sync()
        e1.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        dev_ms = e0.elapsed_time(e1)
        k_ms = runner.env_kernel_ms(args.steps)
        return max(dev_ms, 0.0), wall_ms, runner.env_steps - s0, runner.launches - l0 + 2 * args.steps, k_ms, d2h

    # device-resident arm
    runner = make()
    replay = ReplicatedReplay(torch, dev, capacity=1 << 20, slab_capacity=slab)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms, wall_ms, steps, launches, k_ms, _ = timed(runner, replay)
    clocks = sampler.stop() if rank == 0 else None
    counters = dict(runner.counters)
    replay_size = replay.device_size()
    xbytes = replay.bytes_exchanged / max(1, args.settle + args.warmup + args.steps)
    # end-to-end arm: host-side policy loop + transition records read back to pinned host memory every tick
    hp = HostLoopPolicy(torch, dev, 99 + rank, n, TASK_ADIM[args.task])
    runner2 = make(policy=hp)
    replay2 = ReplicatedReplay(torch, dev, capacity=1 << 20, slab_capacity=slab)
    h_block = torch.zeros(1 + slab, 92, dtype=torch.float32).pin_memory()
    _, wall2_ms, steps2, _, _, d2h_tr = timed(runner2, replay2, h_block)
    e2e_ticks = args.settle + args.warmup + args.steps   # hp counts every tick of timed()
    t = torch.tensor([wall_ms, wall2_ms], dtype=torch.float64, device=dev)
    cnt = torch.tensor([steps, steps2], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    wall_ms, wall2_ms = float(t[0]), float(t[1])
    tot_steps, tot_steps2 = float(cnt[0]), float(cnt[1])
    if rank == 0:
        peak, peak_kind = measured_peaks()
        # fp32 rows in (qpos, qvel, action 8, prev_state 8) + out (qpos, qvel, obs 40, reward / done / pad 4): 816 B for push
        m_ = runner.venv.model
        pad4 = lambda k: (k + 3) // 4 * 4
        bytes_per_env_step = 4 * (2 * pad4(m_.nq) + 2 * pad4(m_.nv) + 8 + 8 + 40 + 4)
        achieved = bytes_per_env_step * n / (k_ms * 1e-3) / 1e9
        cores = os.cpu_count() or 1
        # same protocol as `--impl reference` (one scalar runner per host core, args.cpu_macros macro actions each, rate = env-steps /
        # busy time of the slowest runner); a short sample when other ranks are waiting
        cpu_macros = args.cpu_macros if world == 1 else min(args.cpu_macros, 24)
        cpu_rate, cpu_steps, cpu_busy, _ = cpu_rollout_rate(cores, cpu_macros, 1234, args.max_iter, task=args.task)
        line = {
            "metric": task_metric(args.task), "value": tot_steps / (wall_ms * 1e-3), "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 physics / f32 collision", "data": "synthetic",
            "config": rollout_config(args, n, dev_ms=dev_ms, counters=counters, replay_size=replay_size, xbytes=xbytes, slab=slab),
            "clocks": clocks,
            "e2e": {"value": tot_steps2 / (wall2_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": hp.h2d / e2e_ticks,
                    "d2h_bytes_per_step": hp.d2h / e2e_ticks + d2h_tr / args.steps,
                    "api": "NativeMoPARolloutRunner.tick() with a host-side policy loop (observations D2H, actions H2D, pinned memory) + "
                           "ReplicatedReplay.exchange(); the tick's transition block is read back to pinned host memory"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(args.traffic_file) if args.task == "push" else None,
                         "peak_source": peak_kind, "kernel": "env_step_warp_kernel", "algorithmic_bytes_per_env_step": bytes_per_env_step,
                         "kernel_ms_per_launch": k_ms, "kernel_share_of_step": k_ms * args.steps / max(dev_ms, 1e-9),
                         "second_bound": ncu_pipe(args.traffic_file),
                         "note": "75 substeps per env.step run on chip: the kernel is fp64 latency bound, not HBM bound (see DESIGN.md section 4)"},
            "cpu_baseline": {"value": cpu_rate, "unit": "env-steps/s", "cores": cores, "kind": "port",
                             "sample": "%d macro actions on each of %d scalar runners (%d env-steps, %.1f s)" % (cpu_macros, cores, cpu_steps, cpu_busy)},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def rollout_config(args, n, **extra):
    """The `config` object of the rollout line; both arms print the same workload keys."""
    wl = (WORKLOADS["rollout"] % n).replace("SawyerPushObstacle-v0", TASK_ENV[args.task]).replace("OMEGA", str(TASK_OMEGA[args.task]))
    if args.task == "pusher":
        wl = wl.replace("action_range 0.5, RRT-Connect range 0.1", "action_range 1.0, RRT-Connect range 0.2")
    if args.task == "lift-ik":
        wl = wl.replace("MoPA (", "MoPA-SAC IK (use_ik_target, ik_target grip_site, Cartesian actions through the device IK front end; ").replace("action_range 0.5", "action_range 0.2")
    cfg = {"workload": wl, "reuse_data": True, "max_reuse_data": 30 if args.task == "pusher" else 15, "envs_per_gpu": n,
           "substeps_per_env_step": "100 RK4 mj_steps (400 forward-dynamics evaluations)" if args.task == "pusher" else 75, "max_iter": args.max_iter}
    if args.task == "pusher":
        cfg["baseline_config"] = "BASELINE configs[0] (PusherObstacle-v0 MoPA-SAC; the reference runs it with 1 env on the CPU): vectorised here, the reference arm runs one scalar env per host core"
    if extra:
        cfg.update({
            "settle_ticks": args.settle,
            "settle": "untimed ticks before the W warm-up ticks: every env starts a macro action at tick 0, the planning burst has decayed by tick ~40",
            "l2": "per-tick working set (env state + planner trees) is rewritten every tick; kernels are compute/latency bound",
            "device_ms_per_step": extra["dev_ms"] / args.steps, "counters": extra["counters"],
            "replay": {"records": extra["replay_size"], "slab_rows": extra["slab"], "allgather_bytes_landed_per_tick": extra["xbytes"]}})
    return cfg


def ncu_pipe(name):
    """fp64 / issue utilisation of the env-step kernel from the committed ncu capture (the honest second bound next to HBM)."""
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            d = json.load(f)
        return {k: d[k] for k in ("fp64_pipe_pct", "fp32_issue_frac", "issue_active_pct", "threads_per_instruction", "warps_active_pct") if k in d} or None
    except Exception:
        return None


# ----------------------------------------------------------------------------- GPU arm: validity microbench
def run_validity(args):
    import torch
    import torch.distributed as dist

    from mopa_rl_b200.capi import NativePlanner

    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.task == "assembly":
        raise SystemExit("--workload validity runs on the push (BASELINE config 5) or lift (mesh collider) scene")
    model, ignored, passive, ref = push_setup(args.task)
    planner = NativePlanner(model, passive, ignored, -0.002, 0.1, seed=1234, device=local_rank)
    n = args.queries
    row = ((model.nq + 3) // 4) * 4
    hq = synth_qpos_f32(model, ref, n, 1234 + rank, row)
    h_pinned = torch.from_numpy(hq).pin_memory()
    d_q = h_pinned.to(dev, non_blocking=False)
    d_r = torch.zeros(n, dtype=torch.int32, device=dev)
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
    e2e_steps = max(1, min(args.steps, 5))
    planner.is_valid_host_f32(h_pinned.data_ptr(), row, n, h_words.data_ptr(), 0)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        planner.is_valid_host_f32(h_pinned.data_ptr(), row, n, h_words.data_ptr(), 0)
    barrier()
    e2e_s = time.perf_counter() - t0
    full_rows_e2e_s, e2e_api, e2e_h2d = e2e_s, "mopa_is_valid_host_f32 (pinned host rows of nq floats in, result words out)", n * row * 4
    # The reference's own convention (KinematicPlanner::isValidState): a state is the vector of the planned joints, the passive
    # joints are the planner's.  Applies when the passive entries are the same in every query (push scene: all at qpos0).
    passive_cols = [i for i in range(model.nq) if i not in set(ref)]
    if sorted(ref) == list(ref) and np.all(hq[:, passive_cols] == hq[0, passive_cols]):
        h_active = torch.from_numpy(np.ascontiguousarray(hq[:, ref])).pin_memory()
        base = hq[0, :model.nq].copy()
        h_words_full = h_words.clone()
        planner.is_valid_active_host_f32(h_active.data_ptr(), n, base, h_words.data_ptr(), 0)
        assert torch.equal(h_words, h_words_full), "active-joint and full-row entry points disagree"
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            planner.is_valid_active_host_f32(h_active.data_ptr(), n, base, h_words.data_ptr(), 0)
        barrier()
        e2e_s = time.perf_counter() - t0
        e2e_api = "mopa_is_valid_active_host_f32 (pinned host states of the %d planned joints in - the reference's isValidState convention -, result words out)" % len(ref)
        e2e_h2d = n * len(ref) * 4
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms, e2e_s, full_rows_e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s, full_rows_e2e_s = float(t[0]), float(t[1]), float(t[2])
    words = d_r.cpu().numpy().view(np.uint32)
    assert np.array_equal(words & 1, h_words.numpy().view(np.uint32) & 1), "device-resident and end-to-end paths disagree"
    if rank == 0:
        peak, peak_kind = measured_peaks()
        bytes_per_query = row * 4 + 4
        ms_step = total_ms / args.steps
        achieved = bytes_per_query * n / (float(np.mean(kernel_ms)) * 1e-3) / 1e9
        cores = os.cpu_count() or 1
        ns = min(n, args.cpu_sample)
        rate, dt, ow = oracle_validity_rate(model, ignored, hq[:ns, :model.nq].astype(np.float64), cores)
        mism = int(((ow & 1) != (words[:ns] & 1)).sum())
        print(json.dumps({
            "metric": "state-validity queries/sec (collision-check microbench)", "value": world * n / (ms_step * 1e-3), "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": (WORKLOADS["validity"] % n).replace("SawyerPushObstacle-v0", TASK_ENV[args.task]), "queries_per_gpu": n, "row_bytes": row * 4,
                       "l2": "inputs (%.2f GB per GPU) larger than L2" % (n * row * 4 / 1e9), "valid_fraction": float((words & 1).mean())},
            "clocks": clocks,
            "e2e": {"value": world * n * e2e_steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": e2e_h2d, "d2h_bytes_per_step": n * 4,
                    "api": e2e_api, "full_rows_value": world * n * e2e_steps / full_rows_e2e_s, "full_rows_h2d_bytes_per_step": n * row * 4},
            "gpu_launches": args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (lambda t: None if t is None else int(t * n / 2_000_000))(ncu_traffic("r2_validity_traffic.json") if args.task == "push" else None),
                         "peak_source": peak_kind, "kernel": "is_valid_kernel", "algorithmic_bytes_per_query": bytes_per_query,
                         "second_bound": ncu_pipe("r2_validity_traffic.json") if args.task == "push" else None,
                         "note": "FP32 issue / latency bound, not HBM bound: second_bound.fp32_issue_frac = issue-active x active lanes / 32 from the committed ncu capture"},
            "cpu_baseline": {"value": rate, "unit": "queries/s", "cores": cores, "kind": "port",
                             "sample": "%d of the same queries, one oracle scene per host thread" % ns, "gpu_bit_mismatches": mism}}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10, help="untimed warm-up ticks (after the settle phase)")
    ap.add_argument("--settle", type=int, default=40, help="untimed ticks before the warm-up: lets the start-up planning burst (every env plans at tick 0) decay")
    ap.add_argument("--slab", type=int, default=0, help="rows of the per-tick replay exchange block (0 = envs per GPU, at least 1024)")
    ap.add_argument("--traffic-file", default="r2_envwarp_traffic.json", help="profiles/<file>: dram bytes per launch + pipe utilisation from the ncu capture")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="rollout", choices=["rollout", "validity"])
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU (rollout)")
    ap.add_argument("--task", default="push", choices=["push", "assembly", "lift", "lift-ik", "pusher"],
                    help="rollout scene: push = SawyerPushObstacle-v0 (BASELINE metric, default), assembly = SawyerAssemblyObstacle-v0 (configs[3]), "
                         "lift = SawyerLiftObstacle-v0 (configs[2] scene, joint-space MoPA-SAC; 1024 envs per GPU there), lift-ik = the same scene with the "
                         "MoPA-SAC IK preset of configs[2] (use_ik_target), pusher = PusherObstacle-v0 (configs[0])")
    ap.add_argument("--max-iter", type=int, default=1000, help="RRT-Connect iteration cap (stands in for --timelimit)")
    ap.add_argument("--cpu-macros", type=int, default=150, help="macro actions per scalar runner in the cpu_baseline leg and per step of --impl reference "
                                                                   "(one protocol for both: ~2.5 s of CPU work per runner)")
    ap.add_argument("--queries", type=int, default=10_000_000)
    ap.add_argument("--cpu-sample", type=int, default=2_000_000)
    ap.add_argument("--ref-sample", type=int, default=1_000_000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload == "validity":
        return run_validity(args)
    run_rollout(args)


if __name__ == "__main__":
    main()
