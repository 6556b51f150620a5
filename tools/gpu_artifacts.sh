#!/bin/bash
# Round artifacts on one B200: bench (both arms), ncu launch list + full capture of the scan kernel,
# config-3 chain bench, config-5 sweep.  Outputs under gpurun_out/art/.
O=gpurun_out/art; mkdir -p $O
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_promisc_v7 -s 3 -c 1 -f -o $O/scan_v7_full python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $O/ncu_full_bench.log 2>&1
ncu -i $O/scan_v7_full.ncu-rep --page raw --csv > $O/scan_v7_full_raw.csv 2>/dev/null
ncu -i $O/scan_v7_full.ncu-rep --page source --csv > $O/scan_v7_full_src.csv 2>/dev/null
rm -f $O/scan_v7_full.ncu-rep
python tools/chain_bench.py > $O/chain_config3.json 2> $O/chain.err
python tools/sweep.py > $O/sweep_config5.json 2> $O/sweep.err
cat $O/bench_n1.json; cat $O/bench_reference_arm.json; tail -n 2 $O/*.err
