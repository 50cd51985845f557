#!/bin/bash
mkdir -p gpurun_out
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5"
$T > gpurun_out/bench_2gpu_reach4096.json 2> gpurun_out/bench_2gpu.err; tail -1 gpurun_out/bench_2gpu_reach4096.json | cut -c1-200
$T --task stack --envs 8192 > gpurun_out/bench_2gpu_stack.json 2>> gpurun_out/bench_2gpu.err; tail -1 gpurun_out/bench_2gpu_stack.json | cut -c1-200
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_1of2gpu_reach4096.json 2>> gpurun_out/bench_2gpu.err; cut -c1-120 gpurun_out/bench_1of2gpu_reach4096.json
tail -3 gpurun_out/bench_2gpu.err
