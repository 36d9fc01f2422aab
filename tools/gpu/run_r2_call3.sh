#!/bin/bash
# round 2, call 3: K1r with compact sincos, (e, sign) stash, PE images by TMA, slim barrier waits.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest_gpu.log
timeout 200 python tools/gpu/gpu_time_rgrad.py > $O/k1r_time.txt 2>&1; echo "time rc=$?"; cat $O/k1r_time.txt
timeout 120 python tools/gpu/gpu_clk_rgrad.py > $O/k1r_clk.txt 2>&1; echo "clk rc=$?"; head -20 $O/k1r_clk.txt
timeout 300 python bench.py --no-cpu-baseline --no-gpu-incumbent > $O/bench_train_fp32.json 2> $O/bench_train.err; echo "bench train rc=$?"; cut -c1-200 $O/bench_train_fp32.json; tail -3 $O/bench_train.err
timeout 300 python bench.py --mode infer --no-cpu-baseline --no-gpu-incumbent > $O/bench_infer_fp32.json 2> $O/bench_infer.err; echo "bench infer rc=$?"; cut -c1-200 $O/bench_infer_fp32.json
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"mlp_rgrad" -s 3 -c 1 -o /tmp/prof_k1r python bench.py --mode infer --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-incumbent > $O/ncu_k1r.log 2>&1; echo "ncu k1r rc=$?"
ncu -i /tmp/prof_k1r.ncu-rep --page raw --csv > $O/prof_k1r_raw.csv 2>/dev/null
ncu -i /tmp/prof_k1r.ncu-rep --page source --csv 2>/dev/null | python tools/gpu/ncu_stalls.py > $O/prof_k1r_stalls.txt
ls -la $O | tail -12
