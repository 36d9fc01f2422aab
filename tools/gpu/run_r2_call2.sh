#!/bin/bash
# round 2, call 2: loss-scaled backward, status word, tightened parity tests, bench with the eager-CUDA incumbent.
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tools/gpu/r2_parity_probe.py > $O/r2_parity_probe.log 2>&1; echo "probe rc=$?"; tail -25 $O/r2_parity_probe.log
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $O/pytest_gpu.log
timeout 400 python bench.py > $O/bench_train_fp32.json 2> $O/bench_train.err; echo "bench train rc=$?"; cut -c1-600 $O/bench_train_fp32.json; tail -3 $O/bench_train.err
timeout 300 python bench.py --mode infer --no-cpu-baseline > $O/bench_infer_fp32.json 2> $O/bench_infer.err; echo "bench infer rc=$?"; cut -c1-300 $O/bench_infer_fp32.json
timeout 300 python bench.py --workload c5 --no-cpu-baseline --no-gpu-incumbent --steps 5 > $O/bench_c5.json 2> $O/bench_c5.err; echo "bench c5 rc=$?"; cut -c1-300 $O/bench_c5.json
timeout 300 python bench.py --workload c3 --no-cpu-baseline --no-gpu-incumbent --steps 10 > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench c3 rc=$?"; cut -c1-300 $O/bench_c3.json
timeout 300 python bench.py --workload c2 --no-cpu-baseline --no-gpu-incumbent --steps 10 > $O/bench_c2.json 2> $O/bench_c2.err; echo "bench c2 rc=$?"; cut -c1-300 $O/bench_c2.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cut -c1-400 $O/bench_reference.json
ls -la $O | tail -30
