#!/bin/bash
mkdir -p gpurun_out
HECUDA_CHUNK=256 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lift|tensor|floor|ntt_inv" -s 12 -c 6 -f -o gpurun_out/prof_others python bench.py --steps 1 --warmup 3 --batch 256 --no-e2e --no-cpu-baseline > gpurun_out/ncu_others.log 2>&1
tail -3 gpurun_out/ncu_others.log
