#!/bin/bash
# Light refresh of the round-2 evidence (no `--set full` captures): smoke, bench lines, ncu launch lists (cold = default cache control, and warm).
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/final_smoke.log
timeout 900 python bench.py > $O/final_bench.json 2> $O/final_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/final_bench_reference.json 2> $O/final_bench_reference.err; echo "reference rc=$?"
for w in cifar imagenet; do
  for cc in cold warm; do
    extra=""; [ $cc = warm ] && extra="--cache-control none"
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none $extra -c 640 --csv --log-file $O/final_launches_${w}_${cc}.csv python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-layer-table > $O/final_ncu_${w}_${cc}.log 2>&1
  done
done
ls $O/final_* | head -20
