#!/bin/bash
# 8-GPU session: the metric's batch split over 8 GPUs (strong) + weak, the result gather alone, C5 (100M, 8 shards).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
show() { python -c "
import json,sys; d=json.load(open('$1')); w=d.get('weak') or {}; print('$2 strong', round(d['value']), 'QPS ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), '| weak', round(w.get('value',0)), 'ms', w.get('ms_per_step'))"; }
timeout 600 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 30 > gpurun_out/bench_r2_n8.json 2> gpurun_out/n8.err
show gpurun_out/bench_r2_n8.json N=8 || tail -5 gpurun_out/n8.err
timeout 900 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --workload c5-100Mx96-f16-L2-sharded --steps 5 > gpurun_out/bench_r2_c5.json 2> gpurun_out/c5.err
cut -c1-900 gpurun_out/bench_r2_c5.json; tail -2 gpurun_out/c5.err
timeout 300 $TR --nproc-per-node 4 --master-port 29513 bench.py --gpus 4 --steps 30 > gpurun_out/bench_r2_n4.json 2> gpurun_out/n4.err
show gpurun_out/bench_r2_n4.json N=4 || tail -5 gpurun_out/n4.err
timeout 300 $TR --nproc-per-node 2 --master-port 29514 bench.py --gpus 2 --steps 30 > gpurun_out/bench_r2_n2b.json 2> gpurun_out/n2.err
show gpurun_out/bench_r2_n2b.json N=2 || tail -5 gpurun_out/n2.err
timeout 200 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench_r2_n1_8box.json 2> /dev/null
show gpurun_out/bench_r2_n1_8box.json N=1 || true
timeout 120 $TR --nproc-per-node 8 --master-port 29515 scratch/diag_gather.py 2>/dev/null | grep diag | tee gpurun_out/diag_gather_n8.txt
