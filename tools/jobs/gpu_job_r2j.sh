#!/bin/bash
# round 2, job j (2 GPUs): GPU suite, bench N=1 and N=2 with two-deep pipelining, sharded parity
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2j_pytest.log
tail -5 gpurun_out/r2j_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29561 tools/shard_check.py > gpurun_out/r2j_shard.json 2> gpurun_out/r2j_shard.err; echo "shard_check exit $?"
grep -v "^\s*$" gpurun_out/r2j_shard.err | grep -A8 Traceback | head -20
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err; echo "bench n1 exit $?"
tail -3 gpurun_out/r2j_bench_n1.err
timeout 900 $TR --master-port 29562 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err; echo "bench n2 exit $?"
tail -3 gpurun_out/r2j_bench_n2.err
for f in r2j_bench_n1 r2j_bench_n2; do python -c "
import json; d=json.load(open('gpurun_out/$f.json')); print('$f', round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], (d.get('strong') or {}).get('value'), (d.get('strong') or {}).get('ms_per_step'))"; done
