#!/bin/bash
# compute-sanitizer passes over the parity tests (memcheck: all kernels at small sizes + the tuned NTT sizes;
# racecheck: the shared-memory NTT kernels)
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
K_MEM="ntt_matches_oracle or multiply_matches_oracle_uniform or relinearize_and_modswitch or index_pir or mul_transpose or inner_product_matches or expand_matches or apply_galois_matches or ct_ct_inner_product or power_of_x or plaintext_to_eval"
timeout 1200 $SAN --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -k "$K_MEM" > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck.log
timeout 600 $SAN --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q -k "ntt_matches_oracle or ntt_all_bases or multiply_matches_oracle_uniform" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck.log
