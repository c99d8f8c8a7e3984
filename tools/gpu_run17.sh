#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_seeded.py tests/test_gpu_elementwise.py tests/test_host_mirror.py tests/test_gpu_codec.py -m gpu -q 2>&1 | tail -25
