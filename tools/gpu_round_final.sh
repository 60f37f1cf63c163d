#!/bin/bash
# One gpurun call per round end: GPU parity tests, smoke, bench lines (push default, validity on push + lift, reference arm,
# assembly, lift) and the ncu launch list of the bench command.  Full ncu captures: see tools/gpu_round.sh.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 300 python -m pytest tests -m gpu -x -q > $O/tests_final.log 2>&1; echo "tests exit $?" >> $O/tests_final.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log
timeout 300 python bench.py > $O/bench_final_push.json 2> $O/bench_final_push.err
timeout 200 python bench.py --workload validity > $O/bench_final_validity.json 2> $O/bench_final_validity.err
timeout 200 python bench.py --impl reference > $O/bench_final_reference.json 2> $O/bench_final_reference.err
timeout 200 python bench.py --task assembly --envs 16384 --steps 10 --warmup 3 --cpu-macros 4 > $O/bench_final_assembly.json 2> $O/bench_final_assembly.err
timeout 200 python bench.py --task lift --envs 1024 --steps 30 --warmup 30 --cpu-macros 4 > $O/bench_final_lift_1024.json 2> $O/bench_final_lift.err
timeout 200 python bench.py --task lift --envs 4096 --steps 30 --warmup 30 --cpu-macros 4 > $O/bench_final_lift_4096.json 2>> $O/bench_final_lift.err
timeout 200 python bench.py --workload validity --task lift --steps 20 --warmup 5 > $O/bench_final_validity_lift.json 2> $O/bench_final_validity_lift.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_final.csv python bench.py --steps 3 --warmup 3 --cpu-macros 1 > $O/ncu_bench_final.log 2>&1
tail -3 $O/tests_final.log; tail -2 $O/smoke.log; for f in push validity reference assembly; do cut -c1-200 $O/bench_final_$f.json; echo; done; wc -l $O/launches_final.csv
