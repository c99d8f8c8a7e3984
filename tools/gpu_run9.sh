#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_codec.py -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_codec.log
cat gpurun_out/pytest_codec.log
