#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
: > gpurun_out/sweep12.jsonl
run() { echo "{\"label\": \"$1\"}" >> gpurun_out/sweep12.jsonl; shift; env "$@" >> gpurun_out/sweep12.jsonl 2>> gpurun_out/sweep12.err; }
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
run "reach ls" $B --exec-mode lockstep
run "reach phased" LCR_GROUPS=4 $B --exec-mode phased
run "reach16k ls" $B --exec-mode lockstep --envs 16384
run "reach16k phased" LCR_GROUPS=4 $B --exec-mode phased --envs 16384
