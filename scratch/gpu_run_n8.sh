#!/bin/bash
mkdir -p gpurun_out
run() { # n workload steps port
  DCV_BENCH_WATCHDOG=200 timeout 260 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $4 bench.py --gpus $1 --steps $3 --warmup 5 --workload $2 --no-cpu-baseline > gpurun_out/bench_n$1_$2.log 2>&1
  echo "n$1 $2 rc=$?"; grep '^{' gpurun_out/bench_n$1_$2.log | cut -c1-220
}
run 8 cifar 100 29601
run 8 imagenet 10 29602
run 4 cifar 100 29603
run 4 imagenet 10 29604
