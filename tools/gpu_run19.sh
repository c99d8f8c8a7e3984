#!/bin/bash
# two processes, one per GPU, concurrently: plain vs blocking waits, 4 vs 8 threads
mkdir -p gpurun_out
run2() { # label env threads
  ( env $2 timeout 120 python tools/diag_pir_threads.py plain 0 $3 2>&1 | sed "s/^/[$1 gpu0] /" ) &
  ( env $2 timeout 120 python tools/diag_pir_threads.py plain 1 $3 2>&1 | sed "s/^/[$1 gpu1] /" ); wait
}
{
run2 spin X=1 8
run2 block HECUDA_BLOCKING_SYNC=1 8
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; python -c "import os;print(len(os.sched_getaffinity(0)))"
} > gpurun_out/diag_pir_2proc.log 2>&1
grep -v Warning gpurun_out/diag_pir_2proc.log
