#!/bin/bash
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631"
run() { echo "== $1"; env $1 timeout 200 $TR bench.py --gpus 2 --steps 6 --warmup 3 --no-checks 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['config'].get('grad_reduce_scatter','')[:12])"; }
run "D3_X=0"
run "NCCL_MAX_CTAS=2"
run "NCCL_MAX_CTAS=2 D3_GEMM_SMS=144"
run "NCCL_MAX_CTAS=4 D3_GEMM_SMS=140"
