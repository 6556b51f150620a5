#!/bin/bash
# round 2, job d (1 GPU): ncu --set full per kernel + launch list of the bench
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
cap() { # name target kernel-regex skip
  timeout 600 $NCU -k regex:$3 -s ${4:-0} -c 1 -o gpurun_out/ncu_r2_$1 python tools/ncu_targets.py $2 > gpurun_out/ncu_r2_$1.log 2>&1
  ncu -i gpurun_out/ncu_r2_$1.ncu-rep --page raw --csv > gpurun_out/ncu_r2_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu_r2_$1.ncu-rep --page source --csv > gpurun_out/ncu_r2_$1_src.csv 2>/dev/null
  tail -2 gpurun_out/ncu_r2_$1.log
}
cap decode1 decode1 "decode_kernel" 1
cap decode0 decode0 "decode_kernel" 1
cap tc16 tc16 "decode_kernel" 1
cap sieve sieve "sieve_kernel" 1
cap known known "scan_known_v4" 1
cap k3 k3 "scan_promisc_v7" 1
cap k4 k4 "scan_promisc_v7" 1
NCU_SYMBOLS=2e9 cap k5 k5 "scan_promisc_v7" 1
ls -la gpurun_out/ | grep ncu_r2 | head -40
