#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python tools/elementwise_bench.py --what preprocess > gpurun_out/elementwise_preprocess.log 2>&1; grep bf16 gpurun_out/elementwise_preprocess.log | cut -c1-170
timeout 300 python tools/conv_layer_bench.py --batch 256 --ops fwd,dgrad > gpurun_out/conv_layers.log 2>&1; echo "conv layers rc=$?"; cat gpurun_out/conv_layers.log | tail -7 | cut -c1-230
timeout 300 python bench.py --workload imagenet --no-cpu-baseline > gpurun_out/bench_imagenet.log 2>&1; echo "imagenet rc=$?"; tail -1 gpurun_out/bench_imagenet.log | cut -c1-200
