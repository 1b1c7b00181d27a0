#!/bin/bash
# final single-GPU pass: smoke, full GPU test suite, default bench, ncu launch list + full capture of the attention backward
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 280 python bench.py > gpurun_out/r02_final_bench1.json 2> gpurun_out/r02_final_bench1.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_final_bench1.json').read().strip().splitlines()[-1])
print('bench', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'gemm', d['roofline']['frac'], 'clk', d['clocks'], 'parity', d['parity_check']['rel'], 'cpu', d['cpu_baseline']['value'])"
[ -n "$WITH_NCU" ] && timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 3 -c 1 -o gpurun_out/r02_attn_bwd python tools/bench_attention.py bwd > gpurun_out/ncu_attn_bwd.log 2>&1
[ -n "$WITH_NCU" ] && python tools/ncu_summary.py gpurun_out/r02_attn_bwd.ncu-rep > gpurun_out/r02_ncu_attn_bwd.txt 2>&1; head -5 gpurun_out/r02_ncu_attn_bwd.txt; tail -3 gpurun_out/ncu_attn_bwd.log
exit 0
