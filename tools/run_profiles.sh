#!/bin/bash
# Round-2 evidence captures (1 GPU): labelled GEMM launches, attention forward / backward, launch list of a step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm2sm -s 4 -c 4 -o gpurun_out/r02_ncu_gemm python tools/ncu_gemm_labeled.py > gpurun_out/ncu_gemm.log 2>&1; tail -1 gpurun_out/ncu_gemm.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 16 -c 4 -o gpurun_out/r02_ncu_attn python tools/bench_attention.py all > gpurun_out/ncu_attn.log 2>&1; tail -1 gpurun_out/ncu_attn.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-checks > /dev/null 2>&1; wc -l gpurun_out/r02_launches_final.csv
echo "== ViT-S B=64 (cfg2)"; timeout 200 python bench.py --arch vit_small --steps 8 --warmup 3 --no-cpu-baseline --no-checks 2>/dev/null | cut -c1-330
echo "== ViT-B B=64 (cfg3, 1 GPU)"; timeout 200 python bench.py --arch vit_base --steps 8 --warmup 3 --no-cpu-baseline --no-checks 2>/dev/null | cut -c1-330
echo "== attention isolated"; timeout 60 python tools/bench_attention.py all | grep -E "fwd|bwd"; D3_ATTN_WS=0 timeout 60 python tools/bench_attention.py fwd | grep -E "fwd"
