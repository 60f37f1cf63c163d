#!/bin/bash
# One gpurun call: GPU parity tests, both bench workloads, the reference arm, the ncu launch list of the
# bench command and full captures of the env-step and validity kernels.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
nproc > $O/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests exit $?" >> $O/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log
timeout 900 python bench.py > $O/bench_rollout.json 2> $O/bench_rollout.err; echo "exit $?" >> $O/bench_rollout.err
timeout 600 python bench.py --workload validity > $O/bench_validity.json 2> $O/bench_validity.err
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --cpu-macros 1 > $O/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_step_warp -s 6 -c 1 -f -o $O/envwarp_v4 python tools/time_env.py 4096 --contacts-only > $O/ncu_envwarp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:is_valid_kernel -s 2 -c 1 -f -o $O/validity_v4 python bench.py --workload validity --steps 3 --warmup 3 --queries 2000000 --cpu-sample 1000 > $O/ncu_validity.log 2>&1
timeout 300 python tools/rollout_dynamics.py 4096 400 > $O/dynamics_final.txt 2>&1
timeout 300 python bench.py --task assembly --envs 16384 --steps 10 --warmup 3 --cpu-macros 4 > $O/bench_assembly_16384.json 2> $O/bench_assembly.err
MOPA_PROF_MASKS=0xFE:1:14 timeout 200 python tools/env_prof.py 4096 > $O/env_prof_final.txt 2>&1
ls -la $O
