#!/bin/bash
mkdir -p gpurun_out
DCV_BENCH_WATCHDOG=200 timeout 260 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n2.log 2>&1; echo "n2 cifar rc=$?"; grep '^{' gpurun_out/bench_n2.log | cut -c1-250
DCV_BENCH_WATCHDOG=200 timeout 260 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 --workload imagenet --no-cpu-baseline > gpurun_out/bench_n2_imagenet.log 2>&1; echo "n2 imagenet rc=$?"; grep '^{' gpurun_out/bench_n2_imagenet.log | cut -c1-250
timeout 200 python -m pytest tests/test_dp_gloo.py -x -q 2>&1 | tail -2
