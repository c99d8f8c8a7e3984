#!/bin/bash
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_c3.py 1024 > gpurun_out/bench_c3.log 2> gpurun_out/bench_c3.err
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>> gpurun_out/bench.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_c3.log; tail -2 gpurun_out/bench_c3.err; cat gpurun_out/bench_ref.log | cut -c1-300
