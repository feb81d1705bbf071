#!/bin/bash
# ncu --set full capture of every hot-path kernel (one launch each, from the timed step of a short
# bench run) + the integer-pipe microbenchmark.  Usage on the box: bash tools/gpu_profile.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -x tools/ubench/ubench ]; then tools/ubench/ubench > $OUT/ubench.txt 2>&1; fi
# launch order of `bench.py --steps 1 --warmup 3 --no-e2e --no-cpu`: 4 kernels x (3 warm-up + 1 timed) steps,
# then 7 k_ntt_forward launches: skip the 12 warm-up launches, keep the timed step + 2 NTT launches
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:'k_encode|k_sample|k_encrypt|k_ntt|k_uniform' -s 12 -c 6 -f -o $OUT/hotpath \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --batch ${NCU_BATCH:-16384} > $OUT/ncu_full.log 2>&1
echo "ncu rc=$?"
ls -la $OUT
cat $OUT/ubench.txt
