#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/sweep21.jsonl
run() { echo "{\"label\": \"$1\"}" >> gpurun_out/sweep21.jsonl; shift; env "$@" >> gpurun_out/sweep21.jsonl 2>> gpurun_out/sweep21.err; }
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
run "reach F16" LCR_LS_FLAGS=16 $B
run "reach F17" LCR_LS_FLAGS=17 $B
run "reach F20" LCR_LS_FLAGS=20 $B
run "reach F31" LCR_LS_FLAGS=31 $B
run "reach F23 sort0" LCR_LS_SORT=0 $B
run "lift auto" $B --task lift
run "pp-ee 4096 auto" $B --task pick_place --action-mode ee
timeout 600 compute-sanitizer --tool racecheck python tools/ls_small.py > gpurun_out/racecheck.log 2>&1; tail -3 gpurun_out/racecheck.log
timeout 600 compute-sanitizer --tool memcheck python tools/ls_small.py > gpurun_out/memcheck.log 2>&1; tail -3 gpurun_out/memcheck.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
