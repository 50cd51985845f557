#!/bin/bash
mkdir -p gpurun_out
P=$PWD/gym_lowcostrobot_b200
: > gpurun_out/sweep19.jsonl
run() { echo "{\"label\": \"$1\"}" >> gpurun_out/sweep19.jsonl; shift; env "$@" >> gpurun_out/sweep19.jsonl 2>> gpurun_out/sweep19.err; }
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
run "reach16k dual" $B --envs 16384
run "reach16k nodual" LCR_LIB=$P/liblcrsim_xnodual.so $B --envs 16384
run "reach dual" $B
run "reach nodual" LCR_LIB=$P/liblcrsim_xnodual.so $B
run "reach16k dual again" $B --envs 16384
