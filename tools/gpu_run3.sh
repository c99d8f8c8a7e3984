#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
for c in 29 64 128 256 512; do
  HECUDA_CHUNK=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_chunk$c.log 2>> gpurun_out/bench.err
done
tail -4 gpurun_out/pytest_gpu.log
for c in 29 64 128 256 512; do python - <<PY
import json
d=json.load(open('gpurun_out/bench_chunk$c.log'))
print('chunk $c', round(d['value']), 'mult/s', round(d['ms_per_step'],3),'ms/step', 'ntt/s', round(d['roofline']['ntt_per_s']/1e6,2),'M', 'clk', d['clocks'])
PY
done
tail -3 gpurun_out/bench.err
