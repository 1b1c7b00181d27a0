#!/bin/bash
# 2 ranks on one box: A/B of environment settings, bench.py without the CPU legs; first run includes the fsdp_check
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29631"
run() { echo "== $1 $2"; env $1 timeout 240 $TR bench.py --gpus ${NGPU:-2} --steps 8 --warmup 3 $2 2>gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], 'gemm', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'], d.get('fsdp_check'))" || tail -5 gpurun_out/ab.err; }
for cfg in "$@"; do run "$cfg" "${EXTRA:---no-checks}"; done
