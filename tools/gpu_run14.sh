#!/bin/bash
mkdir -p gpurun_out
PIR_CPU=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    tools/bench_pir.py 1048576 64 8 > gpurun_out/bench_pir_2gpu.log 2> gpurun_out/bench_pir_2gpu.err
echo "pir 2gpu (torchrun) rc=$?"; cut -c1-120 gpurun_out/bench_pir_2gpu.log; grep -o '"n_gpus": [0-9]*, "scaling": "[^"]*", "value": [0-9.]*' gpurun_out/bench_pir_2gpu.log; tail -2 gpurun_out/bench_pir_2gpu.err
echo "--- two independent single-GPU processes at once"
( CUDA_VISIBLE_DEVICES=0 PIR_CPU=0 timeout 200 python tools/bench_pir.py 1048576 64 8 | grep -o '"value": [0-9.]*' | sed 's/^/gpu0 /' ) &
( CUDA_VISIBLE_DEVICES=1 PIR_CPU=0 timeout 200 python tools/bench_pir.py 1048576 64 8 | grep -o '"value": [0-9.]*' | sed 's/^/gpu1 /' ); wait
