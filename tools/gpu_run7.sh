#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multiply or golden" 2>&1 | tail -3 > gpurun_out/pytest_gpu.log
HECUDA_BEHZ_COLS=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multiply or golden" 2>&1 | tail -3 >> gpurun_out/pytest_gpu.log
for c in 2 1; do
  HECUDA_BEHZ_COLS=$c timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_cols$c.log 2>> gpurun_out/bench.err
done
HECUDA_BEHZ_COLS=1 timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --batch 512 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
cat gpurun_out/pytest_gpu.log
for c in 2 1; do python - <<PY
import json
d=json.load(open('gpurun_out/bench_cols$c.log'))
print('cols $c', round(d['value']), 'mult/s', round(d['ms_per_step'],3),'ms/step', d['extra'])
PY
done
tail -3 gpurun_out/bench.err
