#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_ph_sol -s 820 -c 1 -o gpurun_out/prof_phsol python bench.py --steps 2 --warmup 20 --no-cpu-baseline --envs 16384 > gpurun_out/ncu_full_phsol.log 2>&1
tail -2 gpurun_out/ncu_full_phsol.log
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_reach4096.json 2> gpurun_out/bench_reach4096.err; cut -c1-200 gpurun_out/bench_reach4096.json
