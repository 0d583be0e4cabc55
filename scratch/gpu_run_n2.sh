#!/bin/bash
mkdir -p gpurun_out
DCV_BENCH_WATCHDOG=240 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.log 2>&1; echo "n2 cifar rc=$?"; grep '^{' gpurun_out/bench_n2.log | cut -c1-250
DCV_BENCH_WATCHDOG=240 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --workload imagenet > gpurun_out/bench_n2_imagenet.log 2>&1; echo "n2 imagenet rc=$?"; grep '^{' gpurun_out/bench_n2_imagenet.log | cut -c1-250
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 3 --warmup 1 --impl reference > gpurun_out/bench_n2_ref.log 2>&1; echo "n2 ref rc=$?"; grep '^{' gpurun_out/bench_n2_ref.log | cut -c1-250
