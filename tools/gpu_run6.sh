#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
