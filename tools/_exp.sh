#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_golden.py -m gpu -q -k "phased or mode_independent or shards or invariants or packed" 2>&1 | tail -3
: > gpurun_out/sweep35.jsonl
run() { echo "{\"label\": \"$1\"}" >> gpurun_out/sweep35.jsonl; shift; env "$@" >> gpurun_out/sweep35.jsonl 2>> gpurun_out/sweep35.err; }
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
run "reach16k sorted" $B --envs 16384
run "reach16k unsorted" LCR_PH_SORT=0 $B --envs 16384
run "stack8k sorted" $B --task stack --envs 8192
run "stack8k unsorted" LCR_PH_SORT=0 $B --task stack --envs 8192
run "pp8k ee sorted" $B --task pick_place --action-mode ee --envs 8192
run "reach5k sorted" $B --envs 5120
run "reach4k phased sorted" $B --exec-mode phased
run "reach64k sorted" $B --envs 65536 --steps 10
tail -3 gpurun_out/sweep35.err
