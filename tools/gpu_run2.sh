#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
HECUDA_NTT_IMPL=simple timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_simple.log 2>> gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --batch 128 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt_fwd_fast -s 2 -c 2 -f -o gpurun_out/prof_ntt_fwd python bench.py --steps 1 --warmup 3 --batch 64 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -8 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/bench.log; cat gpurun_out/bench_simple.log; tail -3 gpurun_out/bench.err; tail -3 gpurun_out/ncu_full.log
