#!/bin/bash
# round 2, call 5: own weight-gradient contraction kernel (mlp_dw.cu) bring-up + full suite + bench.
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/gpu/gpu_probe_dw.py > $O/dw_probe.txt 2>&1; echo "dw probe rc=$?"; cat $O/dw_probe.txt | cut -c1-400
timeout 900 python -m pytest tests -q -m gpu -s > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "bf16 grad error|passed|failed|FAILED|Error" $O/pytest_gpu.log | head -20
timeout 300 python bench.py --no-cpu-baseline --no-gpu-incumbent > $O/bench_train_fp32.json 2> $O/bench_train.err; echo "bench train rc=$?"; cut -c1-200 $O/bench_train_fp32.json; tail -3 $O/bench_train.err
timeout 300 python bench.py --mode infer --no-cpu-baseline --no-gpu-incumbent > $O/bench_infer_fp32.json 2> $O/bench_infer.err; echo "bench infer rc=$?"; cut -c1-200 $O/bench_infer_fp32.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file $O/launches_train.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-incumbent > $O/ncu_launches.log 2>&1; echo "ncu list rc=$?"
ls -la $O | tail -8
