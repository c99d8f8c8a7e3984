#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/ntt_exp.txt
for d in 0 1 4 5 6; do for p in 256 1024; do HECUDA_NTT_DEBUG=$d python tools/bench_ntt.py $p >> gpurun_out/ntt_exp.txt 2>&1; done; done
cat gpurun_out/ntt_exp.txt
