#!/bin/bash
# Final single-GPU session of the round: tests, the bench line with its CPU baseline and reference arm, launch list,
# ncu --set full of the search kernel, C3 / C4 at their stated sizes.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 240 2>&1 | tail -4
show() { python -c "
import json,sys; d=json.load(open('$1')); r=d['roofline']; c=d.get('cpu_baseline') or {}; print('$2', round(d['value']), 'QPS kernel_ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],3), 'e2e', round(d['e2e']['value']), 'recall', d.get('recall_at_10'), 'cpu', c.get('value'), c.get('ids_identical_to_gpu'), c.get('distances_bit_identical_to_gpu'), 'launches', d['gpu_launches'], d['clocks'])"; }
timeout 400 python bench.py --steps 50 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err
show gpurun_out/bench_r2_n1.json c2 || tail -3 gpurun_out/bench_r2_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_reference.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r2_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r2_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vamana_search_fast -s 6 -c 1 -f -o gpurun_out/prof_r2_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r2_final.log 2>&1
timeout 500 python bench.py --workload c3-1Mx768-f16-IP-w128 --steps 20 > gpurun_out/bench_r2_c3.json 2> gpurun_out/c3.err
show gpurun_out/bench_r2_c3.json c3 || tail -3 gpurun_out/c3.err
timeout 600 python bench.py --workload c4-10Mx96-lvq8-L2-w128 --steps 20 > gpurun_out/bench_r2_c4.json 2> gpurun_out/c4.err
show gpurun_out/bench_r2_c4.json c4 || tail -3 gpurun_out/c4.err
