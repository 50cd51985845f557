#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
: > gpurun_out/sweep6.jsonl
run() { echo "{\"label\": \"$1\"}" >> gpurun_out/sweep6.jsonl; shift; env "$@" >> gpurun_out/sweep6.jsonl 2>> gpurun_out/sweep6.err; }
B="timeout 300 python bench.py --exec-mode lockstep --steps 30 --warmup 5 --no-cpu-baseline"
P=$PWD/gym_lowcostrobot_b200
run "reach W8" LCR_LS_WARPS=8 $B
run "reach W16" LCR_LS_WARPS=16 $B
run "reach W8 fast" LCR_LIB=$P/liblcrsim_fast.so LCR_LS_WARPS=8 $B
run "reach W16 fast" LCR_LIB=$P/liblcrsim_fast.so LCR_LS_WARPS=16 $B
run "reach W16 F31" LCR_LS_WARPS=16 LCR_LS_FLAGS=31 $B
run "reach fused" timeout 300 python bench.py --exec-mode fused --steps 30 --warmup 5 --no-cpu-baseline
run "reach phased" timeout 300 python bench.py --exec-mode phased --steps 30 --warmup 5 --no-cpu-baseline
run "stack W6" LCR_LS_WARPS=6 $B --task stack --envs 8192
run "stack W13" LCR_LS_WARPS=13 $B --task stack --envs 8192
run "push16k W8" LCR_LS_WARPS=8 $B --task push --envs 16384
run "push16k W16" LCR_LS_WARPS=16 $B --task push --envs 16384
run "push16k phased" timeout 300 python bench.py --exec-mode phased --steps 30 --warmup 5 --no-cpu-baseline --task push --envs 16384
run "pp-ee W8" LCR_LS_WARPS=8 $B --task pick_place --action-mode ee --envs 8192
run "reach16k W8" LCR_LS_WARPS=8 $B --envs 16384
LCR_LS_WARPS=16 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step_ls -s 20 -c 1 -o gpurun_out/prof_ls2 python bench.py --exec-mode lockstep --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_ls2.log 2>&1
