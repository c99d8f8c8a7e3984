#!/bin/bash
# 2-GPU pass: sharded PIR (one shard per GPU, NCCL broadcast of relinearization + Galois keys) and the headline bench
mkdir -p gpurun_out
PIR_CPU=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    tools/bench_pir.py 1048576 64 8 > gpurun_out/bench_pir_2gpu.log 2> gpurun_out/bench_pir_2gpu.err
echo "pir 2gpu rc=$?"; cat gpurun_out/bench_pir_2gpu.log; tail -3 gpurun_out/bench_pir_2gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err
echo "bench 2gpu rc=$?"; cut -c1-600 gpurun_out/bench_2gpu.log; tail -3 gpurun_out/bench_2gpu.err
timeout 120 python tools/bench_codec.py 4096 > gpurun_out/bench_codec.log 2>&1; cat gpurun_out/bench_codec.log
