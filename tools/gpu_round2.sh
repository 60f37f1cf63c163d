#!/bin/bash
# Round-2 evidence run (one gpurun call): full GPU test suite, bench lines of every task / workload, the reference arm,
# the ncu launch list of the bench command and a full capture of the env-step kernel inside the rollout.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
nproc > $O/nproc.txt
timeout 1500 python -m pytest tests -m gpu -q -s > $O/r2_gpu_tests.log 2>&1; echo "tests exit $?" >> $O/r2_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1; echo "smoke exit $?" >> $O/r2_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_bench_rollout.json 2> $O/r2_bench_rollout.err
timeout 600 python bench.py --workload validity --steps 20 --warmup 5 > $O/r2_bench_validity.json 2> $O/r2_bench_validity.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2_launches.csv python bench.py --steps 3 --warmup 3 --settle 20 --cpu-macros 1 > $O/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_step_warp -s 30 -c 1 -f -o $O/r2_envwarp python bench.py --steps 3 --warmup 3 --settle 30 --cpu-macros 1 > $O/ncu_envwarp.log 2>&1
ls -la $O
