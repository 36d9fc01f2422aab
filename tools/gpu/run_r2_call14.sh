#!/bin/bash
# round 2, call 14: stash stores after the hand-off in the tangent / dual forward and the reverse sweep.
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python tools/gpu/gpu_time_stages.py > $O/stages_time.txt 2>&1; echo "stages rc=$?"; cut -c1-400 $O/stages_time.txt
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log | cut -c1-600
timeout 400 python bench.py --no-cpu-baseline --no-gpu-incumbent > $O/bench_train_fp32.json 2> $O/bench_train.err; echo "bench train rc=$?"; cut -c1-220 $O/bench_train_fp32.json
