#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
: > gpurun_out/sweep18.jsonl
run() { echo "{\"label\": \"$1\"}" >> gpurun_out/sweep18.jsonl; shift; env "$@" >> gpurun_out/sweep18.jsonl 2>> gpurun_out/sweep18.err; }
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
run "reach auto" $B
run "reach16k auto" $B --envs 16384
run "stack8k auto" $B --task stack --envs 8192
run "pp-ee8k ls" $B --exec-mode lockstep --task pick_place --action-mode ee --envs 8192
timeout 300 python tools/phase_clocks.py ReachCube-v0 4096 25 > gpurun_out/phase_reach8.txt 2>&1; cat gpurun_out/phase_reach8.txt
