#!/bin/bash
# round 2, job i (8 GPUs): bench at N=8 with the fused all-gather (full line) and the copy-engine form (A/B)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2i_bench_n8.json 2> gpurun_out/r2i_bench_n8.err; echo "bench n8 exit $?"
timeout 600 $TR --master-port 29552 bench.py --gpus 8 --steps 20 --warmup 3 --gather ce --no-e2e > gpurun_out/r2i_bench_n8_ce.json 2> gpurun_out/r2i_bench_n8_ce.err; echo "bench n8 ce exit $?"
timeout 600 $TR --master-port 29553 bench.py --gpus 8 --steps 20 --warmup 3 --gather nccl --no-e2e > gpurun_out/r2i_bench_n8_nccl.json 2> gpurun_out/r2i_bench_n8_nccl.err; echo "bench n8 nccl exit $?"
for g in "" _ce _nccl; do python -c "
import json; d=json.load(open('gpurun_out/r2i_bench_n8$g.json')); print('$g', round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'], (d.get('strong') or {}).get('value'), (d.get('strong') or {}).get('ms_per_step'), (d.get('e2e') or {}).get('value'))"; done
