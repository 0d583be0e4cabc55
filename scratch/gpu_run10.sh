#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python tools/elementwise_bench.py --what norm > gpurun_out/elementwise_norm.log 2>&1; echo "elementwise rc=$?"; grep "stats_kernel\|bwd_reduce" gpurun_out/elementwise_norm.log | cut -c1-170
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cifar.log 2>&1; echo "cifar rc=$?"; tail -1 gpurun_out/bench_cifar.log | cut -c1-200
timeout 300 python bench.py --workload imagenet --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_imagenet.log 2>&1; echo "imagenet rc=$?"; tail -1 gpurun_out/bench_imagenet.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'stats_kernel' --launch-skip 2 -c 1 -o gpurun_out/prof_stats -f python tools/elementwise_bench.py --what norm --iters 1 > gpurun_out/ncu_stats.log 2>&1; echo "ncu rc=$?"
