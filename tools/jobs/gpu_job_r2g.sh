#!/bin/bash
# round 2, job g (8 GPUs): bench at N=8 (weak + strong legs), sweep at N=8 against the recorded reference digests, sharded parity
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2g_bench_n8.json 2> gpurun_out/r2g_bench_n8.err; echo "bench n8 exit $?"
tail -3 gpurun_out/r2g_bench_n8.err | cut -c1-300; cut -c1-600 gpurun_out/r2g_bench_n8.json
timeout 900 $TR --master-port 29532 tools/sweep.py > gpurun_out/r02_sweep_config5_n8.json 2> gpurun_out/r2g_sweep.err; echo "sweep n8 exit $?"
tail -3 gpurun_out/r2g_sweep.err | cut -c1-300; tail -2 gpurun_out/r02_sweep_config5_n8.json | cut -c1-500
SHARD_CHECK_SYMBOLS=800000000 timeout 600 $TR --master-port 29533 tools/shard_check.py > gpurun_out/r2g_shard.json 2> gpurun_out/r2g_shard.err; echo "shard_check exit $?"
grep -v "^NCCL" gpurun_out/r2g_shard.json | cut -c1-600
timeout 600 $TR --master-port 29534 bench.py --gpus 8 --steps 20 --warmup 3 --gather nccl --no-e2e --no-extras > gpurun_out/r2g_bench_n8_nccl.json 2> gpurun_out/r2g_bench_n8_nccl.err; echo "bench n8 nccl exit $?"
cut -c1-300 gpurun_out/r2g_bench_n8_nccl.json
