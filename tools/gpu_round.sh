#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch list and one full capture (outputs in gpurun_out/).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
for mode in phased fused; do
  timeout 600 python bench.py --exec-mode $mode --steps 50 --warmup 5 > gpurun_out/bench_reach_$mode.json 2> gpurun_out/bench_reach_$mode.err
done
timeout 600 python bench.py --task push --envs 16384 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_push.json 2> gpurun_out/bench_push.err
timeout 600 python bench.py --task stack --envs 8192 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_stack.json 2> gpurun_out/bench_stack.err
timeout 600 python bench.py --task pick_place --action-mode ee --envs 8192 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pp.json 2> gpurun_out/bench_pp.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_phased.csv python bench.py --exec-mode phased --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_phased.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fused.csv python bench.py --exec-mode fused --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fused.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 4 -c 1 -o gpurun_out/prof_fused python bench.py --exec-mode fused --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_fused.log 2>&1
ls -la gpurun_out
