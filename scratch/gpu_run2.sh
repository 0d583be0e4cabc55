#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cifar.csv python bench.py --no-graph --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cifar.log 2>&1; echo "ncu cifar rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'stats_kernel|bwd_reduce_kernel|bwd_apply_kernel' -c 3 -o gpurun_out/prof_norm -f python tools/elementwise_bench.py --what norm --iters 1 > gpurun_out/ncu_norm.log 2>&1; echo "ncu norm rc=$?"
