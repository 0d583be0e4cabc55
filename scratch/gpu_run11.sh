#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cifar.log 2>&1; echo "cifar rc=$?"; tail -1 gpurun_out/bench_cifar.log | cut -c1-200
DCV_FWD_KC16=1 timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cifar_kc16.log 2>&1; echo "cifar kc16 rc=$?"; tail -1 gpurun_out/bench_cifar_kc16.log | cut -c1-200
timeout 300 python bench.py --workload imagenet --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_imagenet.log 2>&1; echo "imagenet rc=$?"; tail -1 gpurun_out/bench_imagenet.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'conv_fwd_tc_kernel|conv_wgrad_tc_kernel' --launch-skip 2 -c 4 -o gpurun_out/prof_tc_r01 -f python bench.py --workload imagenet --no-graph --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_tc.log 2>&1; echo "ncu tc rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'conv_wgrad_direct_s1_kernel|conv_fwd_direct_s1_kernel' --launch-skip 16 -c 4 -o gpurun_out/prof_direct_r01 -f python bench.py --no-graph --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_direct.log 2>&1; echo "ncu direct rc=$?"
