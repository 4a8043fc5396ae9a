#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 240 2>&1 | tail -6
timeout 200 python scratch/flat_bench.py c2-1Mx96-f32-L2-w128 2>&1 | tail -1 > gpurun_out/flat_r2_c2.json
timeout 200 python scratch/flat_bench.py c3-1Mx768-f16-IP-w128 2>&1 | tail -1 > gpurun_out/flat_r2_c3.json
timeout 200 python scratch/flat_bench.py c2-1Mx96-f32-L2-w128 100 2>&1 | tail -1 > gpurun_out/flat_r2_c2_nq100.json
timeout 200 python scratch/flat_bench.py c2-1Mx96-f32-L2-w128 700 2>&1 | tail -1 > gpurun_out/flat_r2_c2_nq700.json
cut -c1-330 gpurun_out/flat_r2_c2.json gpurun_out/flat_r2_c3.json gpurun_out/flat_r2_c2_nq100.json gpurun_out/flat_r2_c2_nq700.json
timeout 300 python scratch/latency_diag.py 2>/dev/null | tail -1 | tee gpurun_out/latency_diag_r2.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flat_gemm -c 1 -f -o gpurun_out/prof_r2_flat_c3 python scratch/flat_bench.py c3-1Mx768-f16-IP-w128 > gpurun_out/ncu_r2_flat_c3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flat_gemm -c 1 -f -o gpurun_out/prof_r2_flat python scratch/flat_bench.py c2-1Mx96-f32-L2-w128 > gpurun_out/ncu_r2_flat.log 2>&1
