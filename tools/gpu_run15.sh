#!/bin/bash
mkdir -p gpurun_out
{
timeout 120 python tools/diag_pir_threads.py plain 0
timeout 120 python tools/diag_pir_threads.py plain 1
timeout 120 python tools/diag_pir_threads.py torch 0
timeout 120 python tools/diag_pir_threads.py nccl 0
echo "--- two plain processes at once, one per GPU"
timeout 120 python tools/diag_pir_threads.py plain 0 & timeout 120 python tools/diag_pir_threads.py plain 1; wait
} > gpurun_out/diag_pir_threads.log 2>&1
cat gpurun_out/diag_pir_threads.log | grep -v Warning
