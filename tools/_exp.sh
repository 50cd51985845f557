#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_golden.py -m gpu -q -k "phased or mode_independent or golden or shards" 2>&1 | tail -3
: > gpurun_out/sweep26.jsonl
run() { echo "{\"label\": \"$1\"}" >> gpurun_out/sweep26.jsonl; shift; env "$@" >> gpurun_out/sweep26.jsonl 2>> gpurun_out/sweep26.err; }
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
run "reach16k phased" $B --envs 16384
run "stack8k phased" $B --task stack --envs 8192
run "reach8k phased" $B --envs 8192
run "push16k phased" $B --task push --envs 16384
run "reach4k phased" $B --exec-mode phased
run "reach64k phased" $B --envs 65536 --steps 10
