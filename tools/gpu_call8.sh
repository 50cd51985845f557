#!/bin/bash
mkdir -p gpurun_out
P=$PWD/gym_lowcostrobot_b200
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/phase_clocks.py ReachCube-v0 4096 25 > gpurun_out/phase_reach2.txt 2>&1
cat gpurun_out/phase_reach2.txt
: > gpurun_out/sweep8.jsonl
run() { echo "{\"label\": \"$1\"}" >> gpurun_out/sweep8.jsonl; shift; env "$@" >> gpurun_out/sweep8.jsonl 2>> gpurun_out/sweep8.err; }
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --exec-mode lockstep"
run "reach heavy0" LCR_LS_HEAVY=0 $B
run "reach heavy hw4" $B
run "reach heavy hw8" LCR_LS_HEAVY_WARPS=8 $B
run "reach heavy hw2" LCR_LS_HEAVY_WARPS=2 $B
run "reach heavy hw4 W8" LCR_LS_WARPS=8 $B
run "reach heavy hw4 fast" LCR_LIB=$P/liblcrsim_fast.so $B
run "reach16k heavy hw4" $B --envs 16384
run "push16k heavy hw4" $B --task push --envs 16384
run "push16k heavy0" LCR_LS_HEAVY=0 $B --task push --envs 16384
run "stack8k heavy hw4" $B --task stack --envs 8192
run "stack8k heavy0" LCR_LS_HEAVY=0 $B --task stack --envs 8192
run "pp-ee8k heavy hw4" $B --task pick_place --action-mode ee --envs 8192
run "reach64k heavy hw4" $B --envs 65536 --steps 10
