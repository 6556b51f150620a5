#!/bin/bash
# round 2, job h (2 GPUs): fused all-gather parity + bench at N=2 for the three exchange forms
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 tools/shard_check.py > gpurun_out/r2h_shard.json 2> gpurun_out/r2h_shard.err; echo "shard_check exit $?"
grep -v "^\s*$" gpurun_out/r2h_shard.err | grep -A8 Traceback | head -20; grep -v "^NCCL" gpurun_out/r2h_shard.json | cut -c1-900
for g in peer ce nccl; do
  timeout 600 $TR --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 3 --gather $g --no-e2e > gpurun_out/r2h_bench_n2_$g.json 2> gpurun_out/r2h_bench_n2_$g.err; echo "bench $g exit $?"
  python -c "
import json; d=json.load(open('gpurun_out/r2h_bench_n2_$g.json')); print('$g', round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'], (d.get('strong') or {}).get('value'), (d.get('strong') or {}).get('ms_per_step'))"
done
python -m pytest tests/test_gpu_hops.py -m gpu -x -q 2>&1 | tail -3
