#!/bin/bash
# One 8-GPU session: push reduce-scatter check at ViT-L dimensions (both scatter modes), then the headline benches.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
echo "== push check, vector device-scope red" ; D3_FSDP_PUSH_SYS=0 timeout 240 $TR --master-port 29611 tools/check_fsdp_push.py > gpurun_out/r02_push8_vec.log 2>&1; tail -4 gpurun_out/r02_push8_vec.log
echo "== push check, scalar system-scope atomics" ; D3_FSDP_PUSH_SYS=1 timeout 240 $TR --master-port 29612 tools/check_fsdp_push.py > gpurun_out/r02_push8_sys.log 2>&1; tail -4 gpurun_out/r02_push8_sys.log
echo "== bench ViT-L 8 GPUs (nccl reduce-scatter)" ; timeout 300 $TR --master-port 29613 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r02_bench_vitl_8gpu.json 2> gpurun_out/r02_bench_vitl_8gpu.err; tail -c 1500 gpurun_out/r02_bench_vitl_8gpu.json
echo "== bench ViT-g/14 8 GPUs remat B=64" ; timeout 400 $TR --master-port 29614 bench.py --gpus 8 --arch vit_giant2 --patch 14 --local-size 98 --batch 64 --remat --steps 4 --warmup 3 --no-checks > gpurun_out/r02_bench_vitg_8gpu.json 2> gpurun_out/r02_bench_vitg_8gpu.err; tail -c 600 gpurun_out/r02_bench_vitg_8gpu.json
