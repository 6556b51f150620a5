#!/bin/bash
# round 2, first GPU job: parity of the rewritten decode chain + chain timings
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -15 gpurun_out/r2a_pytest.log
timeout 600 python tools/chain_bench.py --blocks 2000 > gpurun_out/r2a_chain.json 2> gpurun_out/r2a_chain.err; echo "chain exit $?"
tail -3 gpurun_out/r2a_chain.err; cat gpurun_out/r2a_chain.json
