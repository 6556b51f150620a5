#!/bin/bash
# quick loop: decode parity + chain timings (+ optional ncu of one target: NCU_TARGET, NCU_KERNEL)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_decode.py tests/test_gpu_chain.py tests/test_sieve.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/chain_bench.py --blocks 2000 > gpurun_out/quick_chain.json 2> gpurun_out/quick_chain.err; echo "chain exit $?"
tail -2 gpurun_out/quick_chain.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/quick_chain.json'))
for k,v in d.items():
    if isinstance(v,dict) and 'ms' in v: print(k[:60], 'ms',round(v['ms'],3),'kernel_ms',round(v.get('decode_kernel_ms',0),3),'Mpkt/s',round(v.get('decode_kernel_packets_per_s',0)/1e6,1))
print(d.get('sieve_parity'), d.get('parity_sample'))
PY
if [ -n "$NCU_TARGET" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:${NCU_KERNEL:-decode_kernel} -s ${NCU_SKIP:-1} -c 1 -o gpurun_out/ncu_q_$NCU_TARGET python tools/ncu_targets.py $NCU_TARGET > gpurun_out/ncu_q_$NCU_TARGET.log 2>&1
  tail -2 gpurun_out/ncu_q_$NCU_TARGET.log
fi
