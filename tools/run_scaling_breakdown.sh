#!/bin/bash
# one box, N GPUs: single-GPU breakdown, then N-rank breakdown (same clocks / same box)
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
timeout 240 python tools/scaling_breakdown.py > gpurun_out/breakdown_1gpu.txt 2> gpurun_out/breakdown_1gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 tools/scaling_breakdown.py > gpurun_out/breakdown_${N}gpu.txt 2> gpurun_out/breakdown_${N}gpu.err
cat gpurun_out/breakdown_1gpu.txt gpurun_out/breakdown_${N}gpu.txt
tail -3 gpurun_out/breakdown_${N}gpu.err
