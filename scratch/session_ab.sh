#!/bin/bash
# A/B session: tests, speculative prefetch modes, host chunk counts.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 240 2>&1 | tail -8
show() { python -c "
import json,sys; d=json.load(open('$1')); r=d['roofline']; print('$2', round(d['value']), 'QPS kernel_ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],3), 'e2e', round(d['e2e']['value']))"; }
for m in 0 1 2; do
  timeout 300 python bench.py --steps 30 --no-cpu-baseline --opt speculative_prefetch=$m > gpurun_out/bench_r2_pf$m.json 2> gpurun_out/pf$m.err
  show gpurun_out/bench_r2_pf$m.json "prefetch[$m]" || tail -3 gpurun_out/pf$m.err
done
for c in 1 2 8; do
  timeout 300 python bench.py --steps 30 --no-cpu-baseline --opt host_chunks=$c > gpurun_out/bench_r2_hc$c.json 2> gpurun_out/hc$c.err
  show gpurun_out/bench_r2_hc$c.json "host_chunks[$c]" || tail -3 gpurun_out/hc$c.err
done
timeout 300 python scratch/batch_sweep.py 2>/dev/null | tail -1 | tee gpurun_out/batch_sweep_r2b.json
for wl in c3-1Mx768-f16-IP-w128; do
  for m in 0 1 2; do
  timeout 400 python bench.py --workload $wl --steps 20 --no-cpu-baseline --opt speculative_prefetch=$m > gpurun_out/bench_r2_c3_pf$m.json 2> gpurun_out/c3pf$m.err
  show gpurun_out/bench_r2_c3_pf$m.json "c3 prefetch[$m]" || tail -3 gpurun_out/c3pf$m.err
  done
done
