#!/bin/bash
# round 2, job m (2 GPUs): parity after the stream clean-up (no copy-engine work on the scan stream) + N=1/N=2 step times
mkdir -p gpurun_out
python -m pytest tests/test_gpu_find_ac.py tests/test_gpu_compat.py tests/test_gpu_chain.py -m gpu -x -q 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29581 tools/shard_check.py > gpurun_out/r2m_shard.json 2> gpurun_out/r2m_shard.err; echo "shard_check exit $?"
grep -v "^\s*$" gpurun_out/r2m_shard.err | grep -A8 Traceback | head -20
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err; echo "bench n1 exit $?"
for g in peer fused; do timeout 900 $TR --master-port 29582 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --gather $g > gpurun_out/r2m_bench_n2_$g.json 2> gpurun_out/r2m_bench_n2_$g.err; echo "bench n2 $g exit $?"; done
for f in r2m_bench_n1 r2m_bench_n2_peer r2m_bench_n2_fused; do python -c "
import json; d=json.load(open('gpurun_out/$f.json')); print('$f', round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'], (d.get('strong') or {}).get('value'), (d.get('strong') or {}).get('ms_per_step'))"; done
