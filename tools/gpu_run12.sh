#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/e2e_sweep.log
for d in 2 3 4; do for st in 4 8 16 32; do
  HECUDA_PIPELINE_DEPTH=$d HECUDA_PIPELINE_STAGES=$st timeout 120 python tools/e2e_sweep.py 1024 >> gpurun_out/e2e_sweep.log 2>&1
done; done
cat gpurun_out/e2e_sweep.log
