#!/bin/bash
mkdir -p gpurun_out
P=$PWD/gym_lowcostrobot_b200
timeout 300 python tools/phase_clocks.py ReachCube-v0 4096 25 > gpurun_out/phase_reach.txt 2>&1
timeout 300 python tools/phase_clocks.py StackTwoCubes-v0 8192 25 > gpurun_out/phase_stack.txt 2>&1
cat gpurun_out/phase_reach.txt
LCR_LIB=$P/liblcrsim_fast.so timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_fast.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_fast.log
tail -5 gpurun_out/pytest_gpu_fast.log
: > gpurun_out/sweep7.jsonl
run() { echo "{\"label\": \"$1\"}" >> gpurun_out/sweep7.jsonl; shift; env "$@" >> gpurun_out/sweep7.jsonl 2>> gpurun_out/sweep7.err; }
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
run "reach ls W16 F16" LCR_LS_FLAGS=16 $B --exec-mode lockstep
run "reach ls W16 F23" $B --exec-mode lockstep
for G in 4 8 16; do
  run "reach phased G$G" LCR_GROUPS=$G $B --exec-mode phased
  run "push16k phased G$G" LCR_GROUPS=$G $B --exec-mode phased --task push --envs 16384
done
run "push16k phased G8 fast" LCR_LIB=$P/liblcrsim_fast.so $B --exec-mode phased --task push --envs 16384
run "stack8k phased" $B --exec-mode phased --task stack --envs 8192
run "pp-ee8k phased" $B --exec-mode phased --task pick_place --action-mode ee --envs 8192
run "reach8k phased" $B --exec-mode phased --envs 8192
run "reach8k ls" $B --exec-mode lockstep --envs 8192
run "reach64k phased" $B --exec-mode phased --envs 65536 --steps 10
