#!/bin/bash
# One gpurun call: GPU parity tests, both bench workloads, the reference arm, kernel timings,
# the ncu launch list of the bench command and one full capture of the env-step kernel.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
nproc > $O/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests exit $?" >> $O/tests.log
timeout 900 python bench.py > $O/bench_rollout.json 2> $O/bench_rollout.err; echo "exit $?" >> $O/bench_rollout.err
timeout 600 python bench.py --workload validity > $O/bench_validity.json 2> $O/bench_validity.err
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 python tools/time_env.py 4096 > $O/time_env.txt 2>&1
timeout 300 python tools/tick_breakdown.py > $O/tick_breakdown.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --cpu-macros 1 > $O/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_step_warp -s 4 -c 1 -f -o $O/envwarp python tools/time_env.py 4096 > $O/ncu_envwarp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 1 -c 1 -f -o $O/plan python tools/tick_breakdown.py > $O/ncu_plan.log 2>&1
ls -la $O
