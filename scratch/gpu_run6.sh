#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python tools/elementwise_bench.py --what im2col > gpurun_out/elementwise_im2col.log 2>&1; echo "elementwise rc=$?"; cat gpurun_out/elementwise_im2col.log | cut -c1-160
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cifar.log 2>&1; echo "cifar rc=$?"; tail -1 gpurun_out/bench_cifar.log | cut -c1-200
timeout 300 python bench.py --workload imagenet --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_imagenet.log 2>&1; echo "imagenet rc=$?"; tail -1 gpurun_out/bench_imagenet.log | cut -c1-200
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_imagenet.csv python bench.py --workload imagenet --no-graph --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_imagenet.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cifar.csv python bench.py --no-graph --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cifar.log 2>&1; echo "ncu cifar rc=$?"
