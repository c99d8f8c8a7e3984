#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_pir.py 1048576 64 8 > gpurun_out/bench_pir_64.log 2> gpurun_out/bench_pir.err
timeout 300 python tools/bench_pir.py 1000000 1 8 > gpurun_out/bench_pir_1.log 2>> gpurun_out/bench_pir.err
cat gpurun_out/bench_pir_64.log gpurun_out/bench_pir_1.log; tail -5 gpurun_out/bench_pir.err
timeout 900 python -m pytest tests -m gpu -q --durations=12 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
