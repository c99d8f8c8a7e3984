#!/bin/bash
# round-end verification pass on one B200: parity suite, smoke, every BASELINE configuration through bench.py
# (+ the reference arm of the headline), wire-format codec.  Outputs under gpurun_out/final_*.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/final_pytest_gpu.txt; cat gpurun_out/final_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.txt 2>&1; tail -2 gpurun_out/final_smoke.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench.err; cut -c1-300 gpurun_out/final_bench_ref.json
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/final_bench.json 2>> gpurun_out/final_bench.err; cut -c1-400 gpurun_out/final_bench.json; tail -2 gpurun_out/final_bench.err
for w in C1 C1-8192 C2-L4 C2-u32 C3 C4 C5; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/final_bench_$w.json 2> gpurun_out/final_bench_$w.err
  echo "$w rc=$? $(cut -c1-200 gpurun_out/final_bench_$w.json)"
done
timeout 120 python tools/bench_codec.py 4096 > gpurun_out/final_bench_codec.json 2>&1; cut -c1-300 gpurun_out/final_bench_codec.json
