#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 240 -x 2>&1 | tail -8
show() { python -c "
import json,sys; d=json.load(open('$1')); r=d['roofline']; print('$2', round(d['value']), 'QPS kernel_ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],3), 'e2e', round(d['e2e']['value']), 'rows', round(r['rows_fetched_per_query'],1))"; }
timeout 300 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench_r2_spec.json 2> gpurun_out/spec.err
show gpurun_out/bench_r2_spec.json "c2" || tail -3 gpurun_out/spec.err
timeout 200 python scratch/batch_sweep.py 2>/dev/null | tail -1 | tee gpurun_out/batch_sweep_r2c.json
timeout 400 python bench.py --workload c3-1Mx768-f16-IP-w128 --steps 20 --no-cpu-baseline > gpurun_out/bench_r2_c3_spec.json 2> gpurun_out/c3spec.err
show gpurun_out/bench_r2_c3_spec.json "c3" || tail -3 gpurun_out/c3spec.err
