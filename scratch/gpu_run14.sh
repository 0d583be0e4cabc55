#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'conv_fwd_tc_halo_kernel|conv_wgrad_tc_kernel' --launch-skip 1 -c 3 -o gpurun_out/prof_halo_r01 -f python tools/conv_layer_bench.py --batch 256 --only s1 --iters 1 > gpurun_out/ncu_halo.log 2>&1; echo "ncu halo rc=$?"
