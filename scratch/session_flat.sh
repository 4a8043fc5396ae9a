#!/bin/bash
# tests, flat search timings + ncu, one GPU's share of the 8-shard C5 job
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 240 2>&1 | tail -8
timeout 200 python scratch/flat_bench.py c2-1Mx96-f32-L2-w128 2>&1 | tail -1 > gpurun_out/flat_r2_c2.json
timeout 200 python scratch/flat_bench.py c3-1Mx768-f16-IP-w128 2>&1 | tail -1 > gpurun_out/flat_r2_c3.json
timeout 200 python scratch/flat_bench.py c2-1Mx96-f32-L2-w128 100 2>&1 | tail -1 > gpurun_out/flat_r2_c2_nq100.json
cut -c1-400 gpurun_out/flat_r2_c2.json gpurun_out/flat_r2_c3.json gpurun_out/flat_r2_c2_nq100.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flat_gemm -c 1 -f -o gpurun_out/prof_r2_flat python scratch/flat_bench.py c2-1Mx96-f32-L2-w128 > gpurun_out/ncu_r2_flat.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flat_gemm -c 1 -f -o gpurun_out/prof_r2_flat_c3 python scratch/flat_bench.py c3-1Mx768-f16-IP-w128 > gpurun_out/ncu_r2_flat_c3.log 2>&1
timeout 200 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench_r2_chunks.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_chunks.json')); r=d['roofline']; print('default', round(d['value']), 'kernel_ms', round(r['kernel_ms'],4), 'e2e', round(d['e2e']['value']), d['ground_truth'])"
timeout 900 python bench.py --workload c5-100Mx96-f16-L2-sharded --shards-total 8 --steps 5 > gpurun_out/bench_r2_c5_share1of8.json 2> gpurun_out/c5share.err
tail -3 gpurun_out/c5share.err; cut -c1-900 gpurun_out/bench_r2_c5_share1of8.json
