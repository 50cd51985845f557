#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cut -c1-300 gpurun_out/bench_default.json
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --task push_loop --envs 16384 > gpurun_out/bench_pushloop16384.json 2>/dev/null; cut -c1-120 gpurun_out/bench_pushloop16384.json
