#!/bin/bash
# ncu --set full of configuration D's uniform-sampler kernels (one launch each) at the 16384-item shard.
# Usage on the box: bash tools/ncu_config_d.sh [tag]; then here: python tools/ncu_config_d_summary.py <tag> <out.json>
TAG=${1:-r02d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for spec in k_uniform_bulk:8 k_uniform_fix:8; do
  K=${spec%%:*}; S=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^(void )?$K" -s $S -c 1 -f \
      -o $OUT/d_$K python tools/run_config_d.py -1 > $OUT/ncu_$K.log 2>&1
  echo "ncu $K rc=$?"
done
ls -la $OUT
