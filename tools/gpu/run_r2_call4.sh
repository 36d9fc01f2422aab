#!/bin/bash
# round 2, call 4: K1r variants (dynamic tiles, LDTM prefetch), per-block load balance, full test-suite.
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/gpu/gpu_time_rgrad.py > $O/k1r_time.txt 2>&1; echo "time rc=$?"; cat $O/k1r_time.txt
timeout 120 python tools/gpu/gpu_clk_rgrad.py > $O/k1r_clk.txt 2>&1; echo "clk rc=$?"; grep -A1 "per-block\|===" $O/k1r_clk.txt | cut -c1-700
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --no-gpu-incumbent > $O/bench_train_fp32.json 2> $O/bench_train.err; echo "bench train rc=$?"; cut -c1-200 $O/bench_train_fp32.json; tail -3 $O/bench_train.err
ls -la $O | tail -8
