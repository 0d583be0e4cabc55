#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/conv_layer_bench.py --batch 256 > gpurun_out/conv_layers.log 2>&1; echo "conv layers rc=$?"; cat gpurun_out/conv_layers.log | tail -7 | cut -c1-330
DCV_TC_NO_HALO=1 timeout 300 python tools/conv_layer_bench.py --batch 256 --ops fwd,dgrad --only s1 > gpurun_out/conv_layers_nohalo.log 2>&1; echo "nohalo rc=$?"; tail -1 gpurun_out/conv_layers_nohalo.log | cut -c1-230
DCV_TC_NO_HALO=1 DCV_TC_NO_RESIDENT=1 timeout 300 python tools/conv_layer_bench.py --batch 256 --ops fwd,dgrad --only s1 > gpurun_out/conv_layers_nohalo_nores.log 2>&1; echo "nohalo nores rc=$?"; tail -1 gpurun_out/conv_layers_nohalo_nores.log | cut -c1-230
timeout 300 python bench.py --workload imagenet --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_imagenet.log 2>&1; echo "imagenet rc=$?"; tail -1 gpurun_out/bench_imagenet.log | cut -c1-200
