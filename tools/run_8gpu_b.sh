#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
echo "== push check at ViT-L dims" ; timeout 200 $TR --master-port 29611 tools/check_fsdp_push.py > gpurun_out/r02_push8_check.log 2>&1; tail -5 gpurun_out/r02_push8_check.log
echo "== bench ViT-L 8 GPUs, push reduce-scatter" ; D3_FSDP_PUSH=1 timeout 300 $TR --master-port 29613 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r02_bench_vitl_8gpu_push.json 2> gpurun_out/r02_bench_vitl_8gpu_push.err; tail -3 gpurun_out/r02_bench_vitl_8gpu_push.err | cut -c1-300; python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_vitl_8gpu_push.json").read().strip().splitlines()[-1])
    print("PUSH8:", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["config"].get("grad_reduce_scatter"), d.get("fsdp_check"))
except Exception as e:
    print("no json:", e)
PY
