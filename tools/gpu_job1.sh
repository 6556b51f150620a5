#!/bin/bash
# developer job: instruction micro-benchmarks, A/B of the scan kernels, ncu source pages
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
tools/ibench > gpurun_out/ibench.txt 2>&1
KBENCH_MODES=${MODES:-v1,v4g,v6,v7,v7f,v7s4,v7s6} python tools/kbench.py --symbols 4000000000 --iters 10 --check > gpurun_out/kbench1.json 2> gpurun_out/kbench1.err
for m in ${NCU_MODES:-v4g v7}; do
  KBENCH_MODES=$m timeout 400 ncu --set full --import-source on --clock-control none -k regex:scan_promisc -c 1 -f -o gpurun_out/ncu_$m python tools/kbench.py --symbols 1000000000 --iters 1 > gpurun_out/ncu_$m.log 2>&1
  ncu -i gpurun_out/ncu_$m.ncu-rep --page raw --csv > gpurun_out/ncu_${m}_raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu_$m.ncu-rep --page source --csv > gpurun_out/ncu_${m}_src.csv 2>/dev/null
done
cat gpurun_out/ibench.txt gpurun_out/kbench1.json
tail -3 gpurun_out/kbench1.err
