#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "im2col_tensor_core" > gpurun_out/pytest_gather.log 2>&1; echo "gather pytest rc=$?"; tail -12 gpurun_out/pytest_gather.log | cut -c1-300
timeout 200 python bench.py --workload imagenet --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_imagenet.log 2>&1; echo "imagenet rc=$?"; tail -1 gpurun_out/bench_imagenet.log | cut -c1-200
DCV_NO_GATHER=1 timeout 200 python bench.py --workload imagenet --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_imagenet_nogather.log 2>&1; echo "imagenet nogather rc=$?"; tail -1 gpurun_out/bench_imagenet_nogather.log | cut -c1-200
