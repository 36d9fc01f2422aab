#!/bin/bash
# round 2, call 6: dynamic tiles in every persistent MLP kernel, bf16 diagnostic, full suite, bench lines.
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/gpu/gpu_time_stages.py > $O/stages_time.txt 2>&1; echo "stages rc=$?"; cat $O/stages_time.txt | cut -c1-400
timeout 200 python tools/gpu/gpu_diag_bf16.py > $O/diag_bf16.txt 2>&1; echo "bf16 rc=$?"; cat $O/diag_bf16.txt
timeout 900 python -m pytest tests -q -m gpu -s > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "bf16 grad error|passed|failed|FAILED|Error" $O/pytest_gpu.log | head -20
timeout 300 python bench.py --no-cpu-baseline --no-gpu-incumbent > $O/bench_train_fp32.json 2> $O/bench_train.err; echo "bench train rc=$?"; cut -c1-200 $O/bench_train_fp32.json; tail -3 $O/bench_train.err
timeout 300 python bench.py --mode infer --no-cpu-baseline --no-gpu-incumbent > $O/bench_infer_fp32.json 2> $O/bench_infer.err; echo "bench infer rc=$?"; cut -c1-200 $O/bench_infer_fp32.json
for wl in c2 c3 c5; do timeout 300 python bench.py --workload $wl --no-cpu-baseline --no-gpu-incumbent --steps 10 > $O/bench_$wl.json 2> $O/bench_$wl.err; echo "bench $wl rc=$?"; cut -c1-160 $O/bench_$wl.json; done
ls -la $O | tail -8
