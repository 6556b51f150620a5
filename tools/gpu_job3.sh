#!/bin/bash
mkdir -p gpurun_out
KBENCH_MODES=v7f timeout 600 ncu --set full --import-source on --clock-control none -k regex:scan_promisc_v7 -c 1 -f -o gpurun_out/ncu_k4 python tools/kbench.py --symbols 1000000000 --iters 1 --k ${K:-4} > gpurun_out/ncu_k4.log 2>&1
ncu -i gpurun_out/ncu_k4.ncu-rep --page raw --csv > gpurun_out/ncu_k4_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_k4.ncu-rep --page source --csv > gpurun_out/ncu_k4_src.csv 2>/dev/null
rm -f gpurun_out/ncu_k4.ncu-rep
tail -2 gpurun_out/ncu_k4.log
