#!/bin/bash
# compute-sanitizer passes over the parity tests (memcheck: every kernel family at small sizes + the tuned NTT sizes;
# racecheck: the kernels that use shared memory -- NTT, AES tables).  Results: profiles/r0N_compute_sanitizer.txt
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
K_MEM="ntt_matches_oracle or multiply_matches_oracle_uniform or relinearize_and_modswitch or index_pir or mul_transpose or inner_product_matches or expand_matches or apply_galois_matches or ct_ct_inner_product or power_of_x or plaintext_to_eval or seeded or wire or codec or elementwise or decrypt or serialize or host_mirror or word32 or u32 or lift_and_floor or auxiliary_base or single_coefficient or key_broadcast_example_single"
timeout 1500 $SAN --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -k "$K_MEM" > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck.log
timeout 900 $SAN --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q -k "ntt_matches_oracle or ntt_all_bases or multiply_matches_oracle_uniform or seeded or wire or ntt_u32" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck.log
