#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload imagenet --no-cpu-baseline > gpurun_out/bench_imagenet.log 2>&1; echo "imagenet rc=$?"; tail -1 gpurun_out/bench_imagenet.log | cut -c1-200
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_cifar.log 2>&1; echo "cifar rc=$?"; tail -1 gpurun_out/bench_cifar.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gather|finalize' -c 60 --csv --log-file gpurun_out/launches_gather.csv python bench.py --workload imagenet --no-graph --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gather2.log 2>&1; echo "ncu rc=$?"; python scratch/launch_summary.py gpurun_out/launches_gather.csv 40 | head -8
