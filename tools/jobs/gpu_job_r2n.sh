#!/bin/bash
# round 2, job n (8 GPUs): N=8 and N=4 lines after the scan stream was freed of copy-engine work
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29591 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2n_bench_n8.json 2> gpurun_out/r2n_bench_n8.err; echo "bench n8 exit $?"
timeout 600 $TR --master-port 29592 bench.py --gpus 8 --steps 20 --warmup 3 --gather fused --no-e2e --no-extras > gpurun_out/r2n_bench_n8_fused.json 2> gpurun_out/r2n_bench_n8_fused.err; echo "bench n8 fused exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29593 bench.py --gpus 4 --steps 20 --warmup 3 --no-e2e > gpurun_out/r2n_bench_n4.json 2> gpurun_out/r2n_bench_n4.err; echo "bench n4 exit $?"
for g in n8 n8_fused n4; do python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_$g.json')); print('$g', round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'], (d.get('strong') or {}).get('value'), (d.get('strong') or {}).get('ms_per_step'), (d.get('e2e') or {}).get('value'))"; done
