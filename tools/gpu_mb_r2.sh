#!/bin/bash
# microbenchmarks of round 2: instruction issue rates and butterfly variants (+ ncu pipe breakdown)
mkdir -p gpurun_out
./tools/mb_r2 > gpurun_out/mb_r2.txt 2>&1
echo "mb rc=$?"
timeout 600 ncu --section ComputeWorkloadAnalysis --section InstructionStats --section SchedulerStats --section WarpStateStats --clock-control none \
   -k regex:"bfly|istream" -f -o gpurun_out/mb_r2_ncu ./tools/mb_r2 > gpurun_out/mb_r2_ncu.log 2>&1
echo "ncu rc=$?"
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv >> gpurun_out/mb_r2.txt
