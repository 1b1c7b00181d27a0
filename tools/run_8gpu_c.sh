#!/bin/bash
# 8 ranks: default bench (with fsdp_check) + kernel timeline of rank 0
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29671"
timeout 300 $TR bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
tail -c 1800 gpurun_out/r02_bench_${N}gpu.json; tail -2 gpurun_out/r02_bench_${N}gpu.err
timeout 200 $TR tools/step_timeline.py 2> gpurun_out/tl${N}.err | head -30
tail -2 gpurun_out/tl${N}.err
