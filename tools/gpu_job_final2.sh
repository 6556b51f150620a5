#!/bin/bash
# second validation job: full GPU suite on the current tree, then compute-sanitizer memcheck over the
# tests of this round's new kernels (decode rewrite incl. forced types, device capture formatter, hops)
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/final2_pytest.txt 2>&1
tail -4 gpurun_out/final2_pytest.txt
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_chain.py tests/test_gpu_decode.py tests/test_gpu_hops.py -m gpu -x -q \
  -k "formatted_on_the_device or forced_type or device_entry_points or hop_winnow" > gpurun_out/final2_memcheck.log 2>&1
echo "memcheck rc=$?" > gpurun_out/final2_memcheck_summary.txt
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/final2_memcheck.log | sort | uniq -c | tail -8 >> gpurun_out/final2_memcheck_summary.txt
cat gpurun_out/final2_memcheck_summary.txt
