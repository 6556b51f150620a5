#!/bin/bash
# third job: ncu re-capture of the winnow kernel (histogram now in shared memory) and of the device capture
# formatter, host-vs-device capture formatting times
O=gpurun_out/artifacts; mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on -f"
cap() { # name target kernel-regex skip count
  timeout 240 $NCU -k regex:$3 -s ${4:-1} -c ${5:-1} -o $O/ncu_$1 python tools/ncu_targets.py $2 > $O/ncu_$1.log 2>&1
  ncu -i $O/ncu_$1.ncu-rep --page raw --csv > $O/ncu_$1_raw.csv 2>/dev/null
  tail -1 $O/ncu_$1.log
}
NCU_REPS=1 cap winnow hops "hop_winnow_kernel" 0
cap capture capture "capture_write_kernel" 1
rm -f $O/*.ncu-rep
python tools/capture_bench.py > gpurun_out/capture_bench.json 2> gpurun_out/capture_bench.err
cat gpurun_out/capture_bench.json; tail -2 gpurun_out/capture_bench.err
