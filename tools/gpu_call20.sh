#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
: > gpurun_out/sweep20.jsonl
run() { echo "{\"label\": \"$1\"}" >> gpurun_out/sweep20.jsonl; shift; env "$@" >> gpurun_out/sweep20.jsonl 2>> gpurun_out/sweep20.err; }
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
run "reach auto" $B
run "reach W14" LCR_LS_WARPS=14 $B
run "reach W12" LCR_LS_WARPS=12 $B
run "reach16k auto" $B --envs 16384
run "stack8k auto" $B --task stack --envs 8192
run "push16k auto" $B --task push --envs 16384
