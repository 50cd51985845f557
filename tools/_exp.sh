#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_lockstep.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log | cut -c1-150
B="timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
$B --task lift > gpurun_out/bench_lift4096.json 2>/dev/null; cut -c1-100 gpurun_out/bench_lift4096.json
$B --task push_loop --action-mode ee --envs 8192 > gpurun_out/bench_pushloop_ee8192.json 2>/dev/null; cut -c1-100 gpurun_out/bench_pushloop_ee8192.json
