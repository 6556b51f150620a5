#!/bin/bash
# round 2, second GPU job (2 GPUs): full -m gpu suite, sharded C ABI parity, bench at N=1 and N=2
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/shard_check.py > gpurun_out/r2b_shard.json 2> gpurun_out/r2b_shard.err; echo "shard_check exit $?"
tail -5 gpurun_out/r2b_shard.err; cat gpurun_out/r2b_shard.json
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo "bench n1 exit $?"
tail -5 gpurun_out/r2b_bench_n1.err; cat gpurun_out/r2b_bench_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err; echo "bench n2 exit $?"
tail -5 gpurun_out/r2b_bench_n2.err; cat gpurun_out/r2b_bench_n2.json
