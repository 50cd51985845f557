#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/diag_hist.py ReachCube-v0 4096 25 > gpurun_out/diag_reach.txt 2>&1; cat gpurun_out/diag_reach.txt
timeout 300 python tools/phase_clocks.py ReachCube-v0 4096 25 > gpurun_out/phase_reach3.txt 2>&1; cat gpurun_out/phase_reach3.txt
: > gpurun_out/sweep9.jsonl
run() { echo "{\"label\": \"$1\"}" >> gpurun_out/sweep9.jsonl; shift; env "$@" >> gpurun_out/sweep9.jsonl 2>> gpurun_out/sweep9.err; }
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
run "reach ls auto" $B --exec-mode lockstep
run "reach ls sorted" LCR_LS_SORT=2 $B --exec-mode lockstep
run "reach phased G4" LCR_GROUPS=4 $B --exec-mode phased
for N in 8192 16384; do
  run "reach$N ls striped" LCR_LS_SORT=1 $B --exec-mode lockstep --envs $N
  run "reach$N ls sorted" LCR_LS_SORT=2 $B --exec-mode lockstep --envs $N
  run "reach$N phased G4" LCR_GROUPS=4 $B --exec-mode phased --envs $N
done
run "reach64k ls sorted" LCR_LS_SORT=2 $B --exec-mode lockstep --envs 65536 --steps 10
run "reach64k phased G4" LCR_GROUPS=4 $B --exec-mode phased --envs 65536 --steps 10
run "push16k ls sorted" LCR_LS_SORT=2 $B --exec-mode lockstep --task push --envs 16384
run "push16k phased G4" LCR_GROUPS=4 $B --exec-mode phased --task push --envs 16384
run "stack8k ls sorted" LCR_LS_SORT=2 $B --exec-mode lockstep --task stack --envs 8192
run "stack8k ls striped" LCR_LS_SORT=1 $B --exec-mode lockstep --task stack --envs 8192
run "stack8k phased G4" LCR_GROUPS=4 $B --exec-mode phased --task stack --envs 8192
run "pp-ee8k ls sorted" LCR_LS_SORT=2 $B --exec-mode lockstep --task pick_place --action-mode ee --envs 8192
run "pp-ee8k phased G4" LCR_GROUPS=4 $B --exec-mode phased --task pick_place --action-mode ee --envs 8192
