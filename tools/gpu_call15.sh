#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/phase_clocks.py ReachCube-v0 4096 25 > gpurun_out/phase_reach7.txt 2>&1; cat gpurun_out/phase_reach7.txt
: > gpurun_out/sweep15.jsonl
run() { echo "{\"label\": \"$1\"}" >> gpurun_out/sweep15.jsonl; shift; env "$@" >> gpurun_out/sweep15.jsonl 2>> gpurun_out/sweep15.err; }
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
run "reach auto" $B
run "reach phased" $B --exec-mode phased
run "reach8k auto" $B --envs 8192
run "reach8k ls" $B --exec-mode lockstep --envs 8192
run "reach16k auto" $B --envs 16384
run "reach16k ls" $B --exec-mode lockstep --envs 16384
run "push16k auto" $B --task push --envs 16384
run "stack8k auto" $B --task stack --envs 8192
run "pp-ee8k auto" $B --task pick_place --action-mode ee --envs 8192
run "pp-ee8k ls" $B --exec-mode lockstep --task pick_place --action-mode ee --envs 8192
run "reach64k auto" $B --envs 65536 --steps 10
