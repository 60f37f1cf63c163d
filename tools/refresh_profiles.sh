#!/bin/bash
# Turns the outputs of tools/gpu_round2_final.sh (gpurun_out/r2f_*) into the tracked summaries under profiles/.
set -e
cd "$(dirname "$0")/.."
O=gpurun_out
T=$(mktemp -d)
(cd $T && cuobjdump -xelf all $OLDPWD/mopa_rl_b200/build/validity_kernel.o >/dev/null 2>&1 && cuobjdump -xelf all $OLDPWD/mopa_rl_b200/build/env_warp.o >/dev/null 2>&1)
python tools/launches_summary.py $O/r2f_launches.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 python bench.py --steps 3 --warmup 3 --settle 20 --cpu-macros 1 (B200, final round-2 build; serialised, cold-cache launch times: shares, not absolutes)" > profiles/r2_rollout_launches_summary.csv
python tools/ncu_summary.py $O/r2f_envwarp.ncu-rep "ncu --set full --clock-control none, env_step_warp_kernel<14,32,24,14,15,14> inside the rollout (bench.py, tick 30 after the settle phase; 4096 envs, SawyerPushObstacle-v0), final round-2 build" > profiles/r2_envwarp_summary.txt
python tools/ncu_summary.py $O/r2f_validity.ncu-rep "ncu --set full --clock-control none, is_valid_kernel<448, false> (SawyerPushObstacle-v0, 2 000 000 queries, fast mode), final round-2 build (reach pruning, window layout, sign-bit masks, warp-scan list emission, sorted refinement queue, rolled loops, one 448-query CTA per SM)" > profiles/r2_validity_v3_summary.txt
python tools/ncu_vk_phases.py $O/r2f_validity.ncu-rep $T/validity_kernel.sm_100a.cubin is_valid_kernelILi448ELb0 30 > profiles/r2_validity_v3_phases.txt
python tools/ncu_by_line.py $O/r2f_validity.ncu-rep $T/validity_kernel.sm_100a.cubin is_valid_kernelILi448ELb0 60 > profiles/r2_validity_v3_by_line.txt 2>/dev/null
python tools/ncu_by_line.py $O/r2f_envwarp.ncu-rep $T/env_warp.sm_100a.cubin env_step_warp_kernelILi14ELi32ELi24ELi14ELi15ELi14 60 > profiles/r2_envwarp_by_line.txt 2>/dev/null
cp $O/r2f_gpu_tests.log profiles/r2_gpu_tests_full_suite.log
python - <<'PY'
import json, re
def metrics(path):
    d = {}
    for l in open(path):
        m = re.match(r"(\S+) = ([0-9.]+)", l)
        if m: d[m.group(1)] = float(m.group(2))
    return d
v = metrics('profiles/r2_validity_v3_summary.txt')
rd, wr = v['dram__bytes_read.sum'] * 1e6, v['dram__bytes_write.sum'] * 1e6
ia, tpi = v['smsp__issue_active.avg.pct_of_peak_sustained_active'], v['smsp__thread_inst_executed_per_inst_executed.ratio']
json.dump({"kernel": "is_valid_kernel<448,false>", "queries": 2000000, "dram_bytes_per_launch": int(rd + wr), "dram_bytes_per_query": (rd + wr) / 2e6,
           "issue_active_pct": ia, "threads_per_instruction": tpi, "fma_pipe_pct": v['sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'],
           "alu_pipe_pct": v['sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'], "warps_active_pct": v['sm__warps_active.avg.pct_of_peak_sustained_active'],
           "fp32_issue_frac": ia / 100 * tpi / 32, "source": "profiles/r2_validity_v3_summary.txt"}, open('profiles/r2_validity_traffic.json', 'w'), indent=1)
e = metrics('profiles/r2_envwarp_summary.txt')
rd, wr = e['dram__bytes_read.sum'] * 1e6, e['dram__bytes_write.sum'] * 1e6
json.dump({"kernel": "env_step_warp_kernel<14,32,24,14,15,14>", "envs": 4096, "dram_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
           "fp64_pipe_pct": e['sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'], "issue_active_pct": e['smsp__issue_active.avg.pct_of_peak_sustained_active'],
           "threads_per_instruction": e['smsp__thread_inst_executed_per_inst_executed.ratio'], "warps_active_pct": e['sm__warps_active.avg.pct_of_peak_sustained_active'],
           "source": "profiles/r2_envwarp_summary.txt"}, open('profiles/r2_envwarp_traffic.json', 'w'), indent=1)
old = [l for l in open('profiles/r2_bench_lines.jsonl') if l.startswith('{')]
multi = [l.strip() for l in old if json.loads(l).get('n_gpus', 1) > 1]
out = [open('gpurun_out/r2f_bench_%s.json' % f).read().strip().splitlines()[-1] for f in ['rollout', 'validity', 'validity_lift', 'reference', 'pusher', 'lift', 'lift_ik', 'assembly']]
open('profiles/r2_bench_lines.jsonl', 'w').write('\n'.join(out + multi) + '\n')
for l in out:
    d = json.loads(l); print(d['metric'][:64], round(d['value']), 'e2e', round(d['e2e']['value']), d.get('roofline', {}).get('kernel_ms_per_launch'))
PY
rm -rf $T
