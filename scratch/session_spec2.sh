#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SVSB200_LIB=$PWD/scratch/variants/libsvsb200_clocks.so timeout 250 python scratch/phase_clocks.py 2>gpurun_out/phase.err | tail -1 | tee gpurun_out/phase_clocks.json
show() { python -c "
import json,sys; d=json.load(open('$1')); r=d['roofline']; print('$2', round(d['value']), 'QPS kernel_ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],3), 'e2e', round(d['e2e']['value']), 'rows', round(r['rows_fetched_per_query'],1))"; }
timeout 300 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench_r2_spec.json 2> gpurun_out/spec.err
show gpurun_out/bench_r2_spec.json "c2" || tail -3 gpurun_out/spec.err
timeout 200 python scratch/batch_sweep.py 2>/dev/null | tail -1 | tee gpurun_out/batch_sweep_r2c.json
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 240 -x 2>&1 | tail -3
