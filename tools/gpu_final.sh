#!/bin/bash
# round-end verification pass on one B200: parity suite, smoke, headline bench (+ reference arm), application benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench_ref.log
timeout 500 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>> gpurun_out/bench.err; cat gpurun_out/bench.log; tail -2 gpurun_out/bench.err
timeout 300 python tools/bench_c3.py 1024 > gpurun_out/bench_c3.log 2> gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.log
timeout 300 python tools/bench_pir.py 1048576 64 8 > gpurun_out/bench_pir_64.log 2> gpurun_out/bench_pir.err; cat gpurun_out/bench_pir_64.log
timeout 300 python tools/bench_pir.py 1000000 1 8 > gpurun_out/bench_pir_1.log 2>> gpurun_out/bench_pir.err; cat gpurun_out/bench_pir_1.log; tail -2 gpurun_out/bench_pir.err
timeout 400 python tools/bench_pnns.py 100000 512 16 > gpurun_out/bench_pnns.log 2> gpurun_out/bench_pnns.err; cat gpurun_out/bench_pnns.log; tail -2 gpurun_out/bench_pnns.err
timeout 120 python tools/bench_codec.py 4096 > gpurun_out/bench_codec.log 2>&1; cat gpurun_out/bench_codec.log
