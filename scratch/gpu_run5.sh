#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python tools/elementwise_bench.py --what norm,im2col > gpurun_out/elementwise_norm.log 2>&1; echo "elementwise rc=$?"; grep "stats_kernel\|im2col" gpurun_out/elementwise_norm.log | cut -c1-160
timeout 200 python tools/conv_layer_bench.py --net cifar --batch 512 > gpurun_out/conv_layers_cifar.log 2>&1; echo "conv cifar rc=$?"; cat gpurun_out/conv_layers_cifar.log | tail -8
DCV_FWD_PX4=1 timeout 200 python tools/conv_layer_bench.py --net cifar --batch 512 > gpurun_out/conv_layers_cifar_px4.log 2>&1; echo "conv cifar px4 rc=$?"; cat gpurun_out/conv_layers_cifar_px4.log | tail -8
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cifar.log 2>&1; echo "cifar rc=$?"; tail -1 gpurun_out/bench_cifar.log | cut -c1-200
DCV_FWD_PX4=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cifar_px4.log 2>&1; echo "cifar px4 rc=$?"; tail -1 gpurun_out/bench_cifar_px4.log | cut -c1-200
timeout 300 python bench.py --workload imagenet --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_imagenet.log 2>&1; echo "imagenet rc=$?"; tail -1 gpurun_out/bench_imagenet.log | cut -c1-200
