#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_cifar.log 2>&1; echo "cifar rc=$?"; tail -1 gpurun_out/bench_cifar.log | cut -c1-200
timeout 300 python bench.py --workload imagenet --no-cpu-baseline > gpurun_out/bench_imagenet.log 2>&1; echo "imagenet rc=$?"; tail -1 gpurun_out/bench_imagenet.log | cut -c1-200
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "reference rc=$?"
