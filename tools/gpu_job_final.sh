#!/bin/bash
# final single-GPU validation of the tree: GPU tests, smoke, classic-call latency, both bench arms
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest.txt 2>&1
tail -3 gpurun_out/final_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.txt 2>&1; tail -1 gpurun_out/final_smoke.txt
python tools/classic_latency.py --reps 20 --find-ac > gpurun_out/final_classic_latency.json 2> gpurun_out/final_classic_latency.err; cat gpurun_out/final_classic_latency.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
cat gpurun_out/final_bench_n1.json
