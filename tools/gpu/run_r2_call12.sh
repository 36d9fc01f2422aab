#!/bin/bash
# round 2, call 12: reverse sweep with the select-light exchange, L2 prefetch of the next layer's stash rows in the
# tangent forward / reverse sweep (A/B), full suite, ncu of the two backward MLP kernels.
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python tools/gpu/gpu_time_stages.py > $O/stages_time.txt 2>&1; echo "stages rc=$?"; cut -c1-400 $O/stages_time.txt
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log | cut -c1-600
timeout 400 python bench.py --no-cpu-baseline --no-gpu-incumbent > $O/bench_train_fp32.json 2> $O/bench_train.err; echo "bench train rc=$?"; cut -c1-220 $O/bench_train_fp32.json
for k in "mlp_kernel<1, 3" "mlp_rev_kernel"; do
  tag=$(echo "$k" | tr -dc 'a-z0-9_')
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:"$k" -s 2 -c 1 -o /tmp/prof_$tag python bench.py --mode train --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-incumbent > $O/ncu_$tag.log 2>&1; echo "ncu $tag rc=$?"
  ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > $O/prof_${tag}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$tag.ncu-rep --page source --csv 2>/dev/null | python tools/gpu/ncu_stalls.py > $O/prof_${tag}_stalls.txt
  head -14 $O/prof_${tag}_stalls.txt | cut -c1-160
done
