#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pir.py -m gpu -q -k "wire" 2>&1 | tail -25
