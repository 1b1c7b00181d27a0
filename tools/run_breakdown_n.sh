#!/bin/bash
# N-rank breakdown only; extra env assignments as further arguments:  run_breakdown_n.sh 2 D3_FSDP_DMA_GATHER=0
cd "$(dirname "$0")/.."
N=${1:-2}; shift
mkdir -p gpurun_out
env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 tools/scaling_breakdown.py 2> gpurun_out/breakdown_${N}gpu.err | grep -v "n=   [12] " | head -${LINES_MAX:-22}
