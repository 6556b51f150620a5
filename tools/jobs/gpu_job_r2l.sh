#!/bin/bash
# round 2, job l (8 GPUs): final N=8 bench line (fused all-gather, two-deep pipelining) + copy-engine A/B
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29571 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2l_bench_n8.json 2> gpurun_out/r2l_bench_n8.err; echo "bench n8 exit $?"
timeout 600 $TR --master-port 29572 bench.py --gpus 8 --steps 20 --warmup 3 --gather ce --no-e2e --no-extras > gpurun_out/r2l_bench_n8_ce.json 2> gpurun_out/r2l_bench_n8_ce.err; echo "bench n8 ce exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus 4 --steps 20 --warmup 3 --no-e2e > gpurun_out/r2l_bench_n4.json 2> gpurun_out/r2l_bench_n4.err; echo "bench n4 exit $?"
for g in n8 n8_ce n4; do python -c "
import json; d=json.load(open('gpurun_out/r2l_bench_$g.json')); print('$g', round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'], (d.get('strong') or {}).get('value'), (d.get('strong') or {}).get('ms_per_step'), (d.get('e2e') or {}).get('value'))"; done
