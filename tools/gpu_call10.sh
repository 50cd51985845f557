#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/phase_clocks.py ReachCube-v0 4096 25 > gpurun_out/phase_reach4.txt 2>&1; cat gpurun_out/phase_reach4.txt
: > gpurun_out/sweep10.jsonl
run() { echo "{\"label\": \"$1\"}" >> gpurun_out/sweep10.jsonl; shift; env "$@" >> gpurun_out/sweep10.jsonl 2>> gpurun_out/sweep10.err; }
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
run "reach ls" $B --exec-mode lockstep
run "reach ls W8" LCR_LS_WARPS=8 $B --exec-mode lockstep
run "reach phased G4" LCR_GROUPS=4 $B --exec-mode phased
run "reach8k ls" $B --exec-mode lockstep --envs 8192
run "reach8k phased G4" LCR_GROUPS=4 $B --exec-mode phased --envs 8192
run "reach16k ls" $B --exec-mode lockstep --envs 16384
run "reach16k phased G4" LCR_GROUPS=4 $B --exec-mode phased --envs 16384
run "push16k ls" $B --exec-mode lockstep --task push --envs 16384
run "push16k phased G4" LCR_GROUPS=4 $B --exec-mode phased --task push --envs 16384
run "stack8k ls" $B --exec-mode lockstep --task stack --envs 8192
run "stack8k phased G4" LCR_GROUPS=4 $B --exec-mode phased --task stack --envs 8192
run "pp-ee8k ls" $B --exec-mode lockstep --task pick_place --action-mode ee --envs 8192
run "pp-ee8k phased G4" LCR_GROUPS=4 $B --exec-mode phased --task pick_place --action-mode ee --envs 8192
run "reach64k ls" $B --exec-mode lockstep --envs 65536 --steps 10
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step_ls -s 20 -c 1 -o gpurun_out/prof_ls3 python bench.py --exec-mode lockstep --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_ls3.log 2>&1
