#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'gather' --launch-skip 1 -c 1 -o gpurun_out/prof_gather -f python bench.py --workload imagenet --no-graph --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gather.log 2>&1; echo "ncu rc=$?"
