#!/bin/bash
# Round artifacts on one B200: -m gpu suite, bench (both arms), ncu launch list of the bench, ncu --set full
# of every kernel on the path.  Outputs under gpurun_out/art/ (tools/mk_profiles.py turns them into profiles/).
O=gpurun_out/art; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/launches_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
cap() { # name target kernel-regex skip count
  timeout 600 $NCU -k regex:$3 -s ${4:-1} -c ${5:-1} -o $O/ncu_$1 python tools/ncu_targets.py $2 > $O/ncu_$1.log 2>&1
  ncu -i $O/ncu_$1.ncu-rep --page raw --csv > $O/ncu_$1_raw.csv 2>/dev/null
  tail -1 $O/ncu_$1.log
}
NCU_SYMBOLS=1e10 cap v7 v7 "scan_promisc_v7" 1
cap known known "scan_known_v4" 1
cap k3 k3 "scan_promisc_v7" 1
cap k4 k4 "scan_promisc_v7" 1
NCU_SYMBOLS=2e9 cap k5 k5 "scan_promisc_v7" 1
NCU_BLOCKS=1000 cap decode1 decode1 "decode_kernel" 1
NCU_BLOCKS=1000 cap decode0 decode0 "decode_kernel" 1
NCU_BLOCKS=1000 cap tc16 tc16 "decode_kernel" 1
NCU_BLOCKS=1000 NCU_REPS=1 cap sieve sieve "sieve_kernel" 0 12
cap hops hops "hop_sequence_kernel" 1
NCU_REPS=1 cap winnow hops "hop_winnow_kernel" 0
NCU_SYMBOLS=1e10 cap slabsort v7 "slab_sort_kernel" 1
rm -f $O/*.ncu-rep
cat $O/bench_n1.json | cut -c1-400; tail -n 2 $O/*.err | cut -c1-300
