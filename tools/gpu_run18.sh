#!/bin/bash
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 300 python -m pytest tests/test_gpu_seeded.py tests/test_gpu_pir.py -m gpu -q -k "seeded or wire" 2>&1 | tail -4
timeout 900 $SAN --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -k "seeded or wire or codec or elementwise or decrypt or serialize or mul_transpose_matrix or host_mirror" > gpurun_out/sanitizer_memcheck2.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck2.log
timeout 600 $SAN --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q -k "seeded or wire" > gpurun_out/sanitizer_racecheck2.log 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/sanitizer_racecheck2.log
PIR_CPU=0 timeout 300 python tools/bench_pir.py 1048576 64 8 > gpurun_out/bench_pir_64b.log 2> gpurun_out/bench_pir.err; cut -c1-700 gpurun_out/bench_pir_64b.log
