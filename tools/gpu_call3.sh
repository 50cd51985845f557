#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
: > gpurun_out/sweep3.jsonl
run() { # label, env assignments..., bench args
  echo "{\"label\": \"$1\"}" >> gpurun_out/sweep3.jsonl; shift
  env "$@" >> gpurun_out/sweep3.jsonl 2>> gpurun_out/sweep3.err
}
B="timeout 300 python bench.py --exec-mode lockstep --steps 30 --warmup 5 --no-cpu-baseline"
run "reach W8 F23" LCR_LS_WARPS=8 LCR_LS_FLAGS=23 $B
run "reach W8 F23 fastmath" LCR_LIB=$PWD/gym_lowcostrobot_b200/liblcrsim_fast.so LCR_LS_WARPS=8 LCR_LS_FLAGS=23 $B
run "reach W16 F23" LCR_LS_WARPS=16 LCR_LS_FLAGS=23 $B
run "reach W5 F23" LCR_LS_WARPS=5 LCR_LS_FLAGS=23 $B
for W in 4 6 13; do
  run "stack W$W F23" LCR_LS_WARPS=$W LCR_LS_FLAGS=23 $B --task stack --envs 8192
done
run "stack W13 F23 fastmath" LCR_LIB=$PWD/gym_lowcostrobot_b200/liblcrsim_fast.so LCR_LS_WARPS=13 LCR_LS_FLAGS=23 $B --task stack --envs 8192
run "push16k W8 F23" LCR_LS_WARPS=8 LCR_LS_FLAGS=23 $B --task push --envs 16384
run "pp-ee W8 F23" LCR_LS_WARPS=8 LCR_LS_FLAGS=23 $B --task pick_place --action-mode ee --envs 8192
run "reach16k W8 F23" LCR_LS_WARPS=8 LCR_LS_FLAGS=23 $B --envs 16384
LCR_LS_WARPS=8 LCR_LS_FLAGS=23 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step_ls -s 20 -c 1 -o gpurun_out/prof_ls python bench.py --exec-mode lockstep --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_ls.log 2>&1
ls -la gpurun_out | tail -5
