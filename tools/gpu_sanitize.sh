#!/bin/bash
# compute-sanitizer over the stage-level and end-to-end parity tests (small batches): memcheck (out-of-bounds /
# misaligned accesses), racecheck (shared-memory hazards — the NTT's __syncwarp()/named-barrier scopes and the
# encode's exchanges on real hardware), synccheck (barrier misuse).  Usage on the box: bash tools/gpu_sanitize.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
SEL="test_ntt or test_encode or test_encrypt_asym or test_encrypt_sym or test_samplers or test_sampler_uniform or test_decrypt or test_gen_public_key or test_sym_partition"
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "$SEL" \
      > $OUT/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $OUT/sanitizer_$tool.log | tail -3
done
