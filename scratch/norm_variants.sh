#!/bin/bash
# A/B of the rows-in-flight / occupancy knobs of the statistics and backward-reduce kernels (rebuilds norm.cu on the GPU box per variant)
for v in "" "-DDCV_SUNR=8 -DDCV_RUNR=8" "-DDCV_SUNR=8 -DDCV_RUNR=8 -DDCV_OCC_CAP=8" "-DDCV_OCC_CAP=8" "-DDCV_SUNR=16 -DDCV_RUNR=8 -DDCV_OCC_CAP=8"; do
  echo "== variant: [$v]"
  DCV_NVCC_FLAGS="$v" python deepcv_b200/csrc/build.py > /dev/null 2>&1
  python tools/elementwise_bench.py --what norm --batch 256 2>/dev/null | grep -E "stats|bwd_reduce" | head -12
done
python deepcv_b200/csrc/build.py > /dev/null 2>&1
