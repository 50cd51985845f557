#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/phase_clocks.py ReachCube-v0 4096 25 > gpurun_out/phase_reach6.txt 2>&1; cat gpurun_out/phase_reach6.txt
: > gpurun_out/sweep14.jsonl
run() { echo "{\"label\": \"$1\"}" >> gpurun_out/sweep14.jsonl; shift; env "$@" >> gpurun_out/sweep14.jsonl 2>> gpurun_out/sweep14.err; }
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
run "reach auto" $B
run "reach ls W8" LCR_LS_WARPS=8 $B --exec-mode lockstep
run "reach phased G4" LCR_GROUPS=4 $B --exec-mode phased
run "reach fused" $B --exec-mode fused
run "reach8k ls striped" LCR_LS_SORT=1 $B --exec-mode lockstep --envs 8192
run "reach8k phased G4" LCR_GROUPS=4 $B --exec-mode phased --envs 8192
run "reach16k ls" $B --exec-mode lockstep --envs 16384
run "reach16k phased G4" LCR_GROUPS=4 $B --exec-mode phased --envs 16384
run "push16k ls" $B --exec-mode lockstep --task push --envs 16384
run "push16k phased G4" LCR_GROUPS=4 $B --exec-mode phased --task push --envs 16384
run "stack8k ls" $B --exec-mode lockstep --task stack --envs 8192
run "stack8k phased G4" LCR_GROUPS=4 $B --exec-mode phased --task stack --envs 8192
run "pp-ee8k ls" $B --exec-mode lockstep --task pick_place --action-mode ee --envs 8192
run "pp-ee8k phased G4" LCR_GROUPS=4 $B --exec-mode phased --task pick_place --action-mode ee --envs 8192
run "lift4k ls" $B --exec-mode lockstep --task lift
run "reach64k ls" $B --exec-mode lockstep --envs 65536 --steps 10
run "reach64k phased" LCR_GROUPS=4 $B --exec-mode phased --envs 65536 --steps 10
