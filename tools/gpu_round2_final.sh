#!/bin/bash
# Final round-2 evidence run (one gpurun call): full GPU test suite, smoke, bench lines of every task / workload, the reference
# arm, the ncu launch list of the bench command and full captures of the env-step kernel (inside the rollout) and of the
# validity kernel.  Everything lands under gpurun_out/r2f_*; the summaries are copied to profiles/ by hand afterwards.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2f_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s > $O/r2f_gpu_tests.log 2>&1; echo "tests exit $?" >> $O/r2f_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2f_smoke.log 2>&1; echo "smoke exit $?" >> $O/r2f_smoke.log
timeout 900 python bench.py > $O/r2f_bench_rollout.json 2> $O/r2f_bench_rollout.err
timeout 600 python bench.py --workload validity > $O/r2f_bench_validity.json 2> $O/r2f_bench_validity.err
timeout 600 python bench.py --workload validity --task lift > $O/r2f_bench_validity_lift.json 2> $O/r2f_bench_validity_lift.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > $O/r2f_bench_reference.json 2> $O/r2f_bench_reference.err
timeout 900 python bench.py --task pusher --steps 20 --warmup 5 > $O/r2f_bench_pusher.json 2> $O/r2f_bench_pusher.err
timeout 900 python bench.py --task lift --envs 1024 --steps 20 --warmup 5 > $O/r2f_bench_lift.json 2> $O/r2f_bench_lift.err
timeout 900 python bench.py --task assembly --envs 16384 --steps 20 --warmup 5 > $O/r2f_bench_assembly.json 2> $O/r2f_bench_assembly.err
timeout 900 python bench.py --task lift-ik --envs 1024 --steps 20 --warmup 5 > $O/r2f_bench_lift_ik.json 2> $O/r2f_bench_lift_ik.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2f_launches.csv python bench.py --steps 3 --warmup 3 --settle 20 --cpu-macros 1 > $O/r2f_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_step_warp -s 30 -c 1 -f -o $O/r2f_envwarp python bench.py --steps 3 --warmup 3 --settle 30 --cpu-macros 1 > $O/r2f_ncu_envwarp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:is_valid_kernel -s 2 -c 1 -f -o $O/r2f_validity python bench.py --workload validity --steps 3 --warmup 3 --queries 2000000 --cpu-sample 1000 > $O/r2f_ncu_validity.log 2>&1
ls -la $O | tail -30
