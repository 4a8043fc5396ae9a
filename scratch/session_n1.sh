#!/bin/bash
# One single-GPU measurement session: tests, bench with CPU baseline, reference arm, launch list, ncu --set full, flat search.
# Every step is bounded by its own timeout so that one stuck step cannot eat the whole call.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 240 2>&1 | tail -8
show() { python -c "
import json,sys; d=json.load(open('$1')); r=d['roofline']; print('$2', round(d['value']), 'QPS kernel_ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],3), 'e2e', round(d['e2e']['value']))"; }
timeout 400 python bench.py --steps 50 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err
show gpurun_out/bench_r2_n1.json "default" || tail -3 gpurun_out/bench_r2_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_n1.json')); c=d['cpu_baseline']; print('recall', d.get('recall_at_10'), 'cpu', c['value'], c.get('ids_identical_to_gpu'), c.get('distances_bit_identical_to_gpu'), 'launches', d['gpu_launches'], d['clocks']); print(d.get('ground_truth'))"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r2_reference.json
timeout 200 python scratch/batch_sweep.py 2>/dev/null | tail -1 | tee gpurun_out/batch_sweep_r2.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r2_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vamana_search_fast -s 3 -c 1 -f -o gpurun_out/prof_r2_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r2_final.log 2>&1
timeout 200 python scratch/flat_bench.py c2-1Mx96-f32-L2-w128 2>&1 | tail -1 > gpurun_out/flat_r2_c2.json
timeout 200 python scratch/flat_bench.py c3-1Mx768-f16-IP-w128 2>&1 | tail -1 > gpurun_out/flat_r2_c3.json
cut -c1-400 gpurun_out/flat_r2_c2.json gpurun_out/flat_r2_c3.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flat_gemm -c 1 -f -o gpurun_out/prof_r2_flat python scratch/flat_bench.py c2-1Mx96-f32-L2-w128 > gpurun_out/ncu_r2_flat.log 2>&1
ls -la gpurun_out/*.ncu-rep
