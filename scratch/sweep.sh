#!/bin/bash
# usage: sweep.sh "<bench args>" ...   prints value e2e frac kernel_ms rows_fetched for each arg set
for a in "$@"; do
  echo "== $a"
  python bench.py --steps 10 --no-cpu-baseline $a 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), round(d['e2e']['value']), round(r['frac'],3), 'kernel_ms', round(r['kernel_ms'],3), 'step_ms', round(d['ms_per_step'],3), r['rows_fetched_per_query'])"
done
