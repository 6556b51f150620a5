#!/bin/bash
# round 2, job c (2 GPUs): -m gpu suite (config-3 chain test, classic routes), sharded C ABI parity at N=2
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest.log
tail -5 gpurun_out/r2c_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/shard_check.py > gpurun_out/r2c_shard.json 2> gpurun_out/r2c_shard.err; echo "shard_check exit $?"
grep -v "^\s*$" gpurun_out/r2c_shard.err | grep -A8 Traceback | head -30; cat gpurun_out/r2c_shard.json
