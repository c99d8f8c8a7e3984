#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --batch 512 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
