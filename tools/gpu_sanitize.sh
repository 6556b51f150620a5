#!/bin/bash
# compute-sanitizer over the GPU parity tests that exercise the bulk kernels and the sieve (small inputs)
O=gpurun_out/sanitize; mkdir -p $O
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_find_ac.py tests/test_sieve.py -m gpu -x -q \
    -k "edge_lengths or barker_dense or alignment or sieve_matches or golden_fixture_cases" > $O/$tool.log 2>&1
  echo "$tool rc=$?" >> $O/summary.txt
  grep -E "ERROR SUMMARY|passed|failed" $O/$tool.log | tail -3 >> $O/summary.txt
done
cat $O/summary.txt
# this round's new kernels (decode rewrite with forced types, device capture formatter, winnow): tools/gpu_job_final2.sh
python tools/kbench.py --symbols 4000000000 --iters 5 --lap 0x9e8b33 --check 2>&1 | tail -1
