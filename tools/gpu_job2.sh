#!/bin/bash
# developer job: A/B of the scan kernels (+ optional ncu source pages)
mkdir -p gpurun_out
KBENCH_MODES=${MODES:-v1,v4g,v7f} python tools/kbench.py --symbols ${SYMS:-4000000000} --iters 10 --check > gpurun_out/kbench2.json 2> gpurun_out/kbench2.err
for m in ${NCU_MODES}; do
  KBENCH_MODES=$m timeout 400 ncu --set full --import-source on --clock-control none -k regex:scan_promisc -c 1 -f -o gpurun_out/ncu_$m python tools/kbench.py --symbols ${NCU_SYMS:-4000000000} --iters 1 > gpurun_out/ncu_$m.log 2>&1
  ncu -i gpurun_out/ncu_$m.ncu-rep --page raw --csv > gpurun_out/ncu_${m}_raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu_$m.ncu-rep --page source --csv > gpurun_out/ncu_${m}_src.csv 2>/dev/null
  rm -f gpurun_out/ncu_$m.ncu-rep
done
python - <<'PY'
import json
d=json.load(open('gpurun_out/kbench2.json'))
for k,v in d.items():
    if isinstance(v,dict): print(k, round(v['GBps'],1), v['hits'], v.get('sha'))
print('match', d.get('match'))
PY
tail -3 gpurun_out/kbench2.err
