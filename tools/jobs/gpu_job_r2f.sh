#!/bin/bash
# round 2, job f (1 GPU): GPU suite, sweep at N=1 with the reference over the whole stream per cell, bench N=1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_pytest.log
tail -4 gpurun_out/r2f_pytest.log
timeout 1500 python tools/sweep.py --cpu-digests gpurun_out/r02_sweep_reference_digests.json > gpurun_out/r02_sweep_config5_n1.json 2> gpurun_out/r2f_sweep.err; echo "sweep exit $?"
tail -3 gpurun_out/r2f_sweep.err; tail -2 gpurun_out/r02_sweep_config5_n1.json | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench exit $?"
tail -3 gpurun_out/r2f_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "ref arm exit $?"
cut -c1-300 gpurun_out/r2f_bench_ref.json
