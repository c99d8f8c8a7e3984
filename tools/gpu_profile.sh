#!/bin/bash
# round-end profiling pass: launch list of the bench command + one full capture of the dominant kernel
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ntt_fwd_fast -s 4 -c 2 -f -o gpurun_out/prof_ntt_fwd \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:"lift_kernel|tensor_kernel|floor_kernel|ntt_inv_fast|ks_mac" -s 8 -c 7 -f -o gpurun_out/prof_others \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_others.log 2>&1
echo "others rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
