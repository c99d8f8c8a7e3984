#!/bin/bash
# every BASELINE configuration through bench.py (one GPU), JSON lines under gpurun_out/
mkdir -p gpurun_out
for w in ${WORKLOADS:-C1 C1-8192 C2 C3 C4 C5}; do
  timeout 900 python bench.py --workload $w --steps ${STEPS:-5} --warmup 3 ${BENCH_FLAGS:-} > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "$w rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1])
    r=d.get('roofline') or {}
    e=d.get('e2e') or {}
    print('value %.4g %s  ms/step %.4g  roofline frac %.3f  e2e %s  cpu %s  extra_err %s' % (d['value'], d['unit'], d['ms_per_step'], r.get('frac', float('nan')), e.get('value'), (d.get('cpu_baseline') or {}).get('value'), (d.get('extra') or {}).get('error')))
except Exception as exc:
    print('no line:', exc)
PY
)"
  tail -3 gpurun_out/bench_$w.err
done
