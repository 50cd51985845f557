#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "lockstep" > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
: > gpurun_out/sweep5.jsonl
run() { echo "{\"label\": \"$1\"}" >> gpurun_out/sweep5.jsonl; shift; env "$@" >> gpurun_out/sweep5.jsonl 2>> gpurun_out/sweep5.err; }
B="timeout 300 python bench.py --exec-mode lockstep --steps 30 --warmup 5 --no-cpu-baseline"
P=$PWD/gym_lowcostrobot_b200
run "reach W8 sort0" LCR_LS_WARPS=8 LCR_LS_SORT=0 $B
run "reach W8 stripe" LCR_LS_WARPS=8 $B
run "reach W8 sort0 nochol" LCR_LIB=$P/liblcrsim_xnochol.so LCR_LS_WARPS=8 LCR_LS_SORT=0 $B
run "reach W8 stripe unroll8" LCR_LIB=$P/liblcrsim_xunroll8.so LCR_LS_WARPS=8 $B
run "reach W8 stripe fast" LCR_LIB=$P/liblcrsim_fast.so LCR_LS_WARPS=8 $B
run "reach W16 stripe" LCR_LS_WARPS=16 $B
run "reach W4 stripe" LCR_LS_WARPS=4 $B
run "reach W6 stripe carve86" LCR_LS_WARPS=6 LCR_LS_CARVEOUT=86 $B
run "reach W6 stripe" LCR_LS_WARPS=6 $B
run "reach W8 sort0 again" LCR_LS_WARPS=8 LCR_LS_SORT=0 $B
run "stack W6 stripe" LCR_LS_WARPS=6 $B --task stack --envs 8192
run "stack W6 sort0" LCR_LS_WARPS=6 LCR_LS_SORT=0 $B --task stack --envs 8192
run "stack W13 stripe" LCR_LS_WARPS=13 $B --task stack --envs 8192
run "push16k W8 stripe" LCR_LS_WARPS=8 $B --task push --envs 16384
run "push16k W8 sort0" LCR_LS_WARPS=8 LCR_LS_SORT=0 $B --task push --envs 16384
run "pp-ee W8 stripe" LCR_LS_WARPS=8 $B --task pick_place --action-mode ee --envs 8192
run "reach16k W8 stripe" LCR_LS_WARPS=8 $B --envs 16384
