#!/bin/bash
# ncu --set full capture of every hot-path kernel (one launch each, from the timed step of a short
# bench run) + the integer-pipe microbenchmark.  Usage on the box: bash tools/gpu_profile.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -x tools/ubench/ubench ]; then tools/ubench/ubench > $OUT/ubench.txt 2>&1; fi
# `bench.py --steps 1 --warmup 3 --no-e2e --no-cpu` launches every step kernel 3 (warm-up) + 1 (timed) + 1
# (per-kernel pass) times; key generation adds one small k_sample_cbd and ntt(s) one small k_ntt_forward
# in front.  One capture per kernel, skipping to the timed step's launch.
for spec in k_encode:3 k_sample_ternary:3 k_sample_cbd:4 k_encrypt_asym:3 k_ntt_forward:3; do
  K=${spec%%:*}; S=${spec##*:}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^(void )?$K" -s $S -c 1 -f \
      -o $OUT/hot_$K python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-other --batch ${NCU_BATCH:-65536} \
      > $OUT/ncu_$K.log 2>&1
  echo "ncu $K rc=$?"
done
ls -la $OUT
