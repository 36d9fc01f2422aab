#!/bin/bash
# round 2, call 8: inference without the training stash, fewer ATen launches, RenderingNetwork op; ncu of K1r in both modes.
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/gpu/gpu_time_stages.py > $O/stages_time.txt 2>&1; cat $O/stages_time.txt | cut -c1-300
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --no-gpu-incumbent > $O/bench_train_fp32.json 2> $O/bench_train.err; echo "bench train rc=$?"; cut -c1-200 $O/bench_train_fp32.json; tail -3 $O/bench_train.err
timeout 300 python bench.py --mode infer --no-cpu-baseline --no-gpu-incumbent > $O/bench_infer_fp32.json 2> $O/bench_infer.err; echo "bench infer rc=$?"; cut -c1-200 $O/bench_infer_fp32.json
for wl in c2 c3; do timeout 300 python bench.py --workload $wl --no-cpu-baseline --no-gpu-incumbent --steps 10 > $O/bench_$wl.json 2> $O/bench_$wl.err; echo "bench $wl rc=$?"; cut -c1-160 $O/bench_$wl.json; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file $O/launches_train.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-incumbent > $O/ncu_launches.log 2>&1; echo "ncu list rc=$?"
for mode in infer train; do
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:"mlp_rgrad" -s 3 -c 1 -o /tmp/prof_k1r_$mode python bench.py --mode $mode --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-incumbent > $O/ncu_k1r_$mode.log 2>&1; echo "ncu k1r $mode rc=$?"
  ncu -i /tmp/prof_k1r_$mode.ncu-rep --page raw --csv > $O/prof_k1r_${mode}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_k1r_$mode.ncu-rep --page source --csv 2>/dev/null | python tools/gpu/ncu_stalls.py > $O/prof_k1r_${mode}_stalls.txt
done
ls -la $O | tail -12
