#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/bench_pnns.py 100000 512 16 > gpurun_out/bench_pnns.log 2> gpurun_out/bench_pnns.err
cat gpurun_out/bench_pnns.log; tail -5 gpurun_out/bench_pnns.err
timeout 300 python tools/bench_pir.py 1048576 64 8 > gpurun_out/bench_pir_64.log 2> gpurun_out/bench_pir.err
cat gpurun_out/bench_pir_64.log; tail -3 gpurun_out/bench_pir.err
