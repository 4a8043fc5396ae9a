#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_flat.py -m gpu -q --timeout 120 2>&1 | tail -4
timeout 200 python scratch/flat_bench.py c2-1Mx96-f32-L2-w128 2>&1 | tail -1 > gpurun_out/flat_r2_c2.json
timeout 200 python scratch/flat_bench.py c3-1Mx768-f16-IP-w128 2>&1 | tail -1 > gpurun_out/flat_r2_c3.json
cut -c1-330 gpurun_out/flat_r2_c2.json gpurun_out/flat_r2_c3.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flat_gemm -c 1 -f -o gpurun_out/prof_r2_flat_c3 python scratch/flat_bench.py c3-1Mx768-f16-IP-w128 > gpurun_out/ncu_r2_flat_c3.log 2>&1
ncu -i gpurun_out/prof_r2_flat_c3.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys; r=list(csv.reader(sys.stdin)); h=r[0]; v=r[2]
for k in ('dram__bytes_read.sum','gpu__time_duration.sum','lts__t_sector_hit_rate.pct'): print(k, v[h.index(k)])"
