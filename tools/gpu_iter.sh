#!/bin/bash
# one development iteration on the GPU box: parity of the core path, a short bench, one full capture of the NTT kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/iter_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/iter_pytest.log)"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/iter_bench.json 2> gpurun_out/iter_bench.err
echo "bench rc=$?"; cat gpurun_out/iter_bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'ntt/s',d['roofline']['ntt_per_s'],'frac',d['roofline']['frac'], d['extra'].get('error'))"
if [ "$1" != "noprof" ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ntt_rows_kernel -s 6 -c 2 -f -o gpurun_out/iter_ntt \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/iter_ncu.log 2>&1
echo "ncu rc=$?"
fi
