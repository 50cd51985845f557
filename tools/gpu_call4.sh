#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
: > gpurun_out/sweep4.jsonl
run() { echo "{\"label\": \"$1\"}" >> gpurun_out/sweep4.jsonl; shift; env "$@" >> gpurun_out/sweep4.jsonl 2>> gpurun_out/sweep4.err; }
B="timeout 300 python bench.py --exec-mode lockstep --steps 30 --warmup 5 --no-cpu-baseline"
for S in 1 0; do for W in 8 4 2; do
  run "reach W$W sort$S" LCR_LS_WARPS=$W LCR_LS_SORT=$S $B
done; done
run "reach W16 sort1" LCR_LS_WARPS=16 LCR_LS_SORT=1 $B
run "reach W8 sort1 F16" LCR_LS_WARPS=8 LCR_LS_FLAGS=16 $B
run "reach W4 sort1 F16" LCR_LS_WARPS=4 LCR_LS_FLAGS=16 $B
run "reach W4 sort1 fast" LCR_LIB=$PWD/gym_lowcostrobot_b200/liblcrsim_fast.so LCR_LS_WARPS=4 $B
for S in 1 0; do for W in 6 4 3; do
  run "stack W$W sort$S" LCR_LS_WARPS=$W LCR_LS_SORT=$S $B --task stack --envs 8192
done; done
run "push16k W8 sort1" LCR_LS_WARPS=8 $B --task push --envs 16384
run "push16k W4 sort1" LCR_LS_WARPS=4 $B --task push --envs 16384
run "pp-ee W4 sort1" LCR_LS_WARPS=4 $B --task pick_place --action-mode ee --envs 8192
run "reach16k W4 sort1" LCR_LS_WARPS=4 $B --envs 16384
run "reach64k W4 sort1" LCR_LS_WARPS=4 $B --envs 65536 --steps 10
