#!/bin/bash
# Round-2 evidence run (one B200): smoke, bench lines, ncu launch lists (default = cold caches, and warm) and one `--set full` capture per workload.
# Outputs under gpurun_out/ (the .ncu-rep files are exported to csv on the box and removed: gpurun brings back at most 64 MiB).
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/final_smoke.log
timeout 900 python bench.py > $O/final_bench.json 2> $O/final_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/final_bench_reference.json 2> $O/final_bench_reference.err; echo "reference rc=$?"
for w in cifar imagenet; do
  for cc in cold warm; do
    extra=""; [ $cc = warm ] && extra="--cache-control none"
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none $extra -c 700 --csv --log-file $O/final_launches_${w}_${cc}.csv python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-layer-table > $O/final_ncu_${w}_${cc}.log 2>&1
  done
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sc_" --launch-skip 60 --launch-count 14 -f -o $O/final_full_cifar python bench.py --workload cifar --steps 2 --warmup 3 --no-cpu-baseline > $O/final_full_cifar.log 2>&1
ncu -i $O/final_full_cifar.ncu-rep --page raw --csv > $O/final_full_cifar_raw.csv 2>/dev/null
ncu -i $O/final_full_cifar.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:"sc_bwd_kernel" --launch-count 1 > $O/final_full_cifar_scbwd_src.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pairs_kernel|halo_kernel" --launch-skip 8 --launch-count 4 -f -o $O/final_full_imagenet python bench.py --workload imagenet --steps 2 --warmup 3 --no-cpu-baseline --no-layer-table > $O/final_full_imagenet.log 2>&1
ncu -i $O/final_full_imagenet.ncu-rep --page raw --csv > $O/final_full_imagenet_raw.csv 2>/dev/null
rm -f $O/*.ncu-rep
du -sh $O; ls $O | head -40
