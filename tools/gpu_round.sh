#!/bin/bash
# One GPU-box visit: parity tests, smoke, the default bench line (all BASELINE configs), the reference arm, the ncu launch list
# of the timed steps and one full capture of the dominant kernel (outputs in gpurun_out/<tag>_*; copy what is to be
# judged into profiles/).
# usage (from the repo root):  gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r02a [quick]'
TAG=${1:-r02}
O=gpurun_out/$TAG
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > ${O}_smi.txt
timeout 900 python -m pytest tests -m gpu -q > ${O}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1
timeout 600 python bench.py > ${O}_bench_default.json 2> ${O}_bench_default.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > ${O}_bench_ref.json 2> ${O}_bench_ref.err
[ "$2" = quick ] && exit 0
# ncu passes: launches of the timed steps only (bench.py brackets them with cudaProfilerStart / Stop); LCR_GRAPH=0 = the same kernels
# launched one by one instead of replayed as one graph
export LCR_GRAPH=0
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_active.avg,sm__cycles_elapsed.max --clock-control none -c 900 --csv --log-file ${O}_launches_push16384.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > ${O}_ncu_list.log 2>&1
for K in k_ph_sol k_ph_job; do
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$K -s 45 -c 1 -o ${O}_prof_$K python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs > ${O}_ncu_full_$K.log 2>&1
done
unset LCR_GRAPH
timeout 200 python tools/ph_timeline.py push joint 16384 ${O}_timeline_push16384.json 2 > ${O}_timeline.log 2>&1
timeout 120 python tools/img_time.py > ${O}_img_time.log 2>&1
ls -la gpurun_out | tail -20
