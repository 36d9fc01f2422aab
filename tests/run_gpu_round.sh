#!/bin/bash
# One gpurun call: smoke, tests, bench lines, ncu launch list + full captures of the top kernels.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_train_fp32.json 2> gpurun_out/bench_train.err; echo "bench train rc=$?"; cut -c1-400 gpurun_out/bench_train_fp32.json
timeout 300 python bench.py --steps 10 --warmup 3 --mode infer > gpurun_out/bench_infer_fp32.json 2> gpurun_out/bench_infer.err; echo "bench infer rc=$?"; cut -c1-300 gpurun_out/bench_infer_fp32.json
timeout 200 python bench.py --steps 10 --warmup 3 --mode infer --precision fp16 --no-cpu-baseline > gpurun_out/bench_infer_fp16.json 2>/dev/null; cut -c1-300 gpurun_out/bench_infer_fp16.json
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 160 --csv --log-file gpurun_out/launches_train.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"mlp_kernel|mlp_rev" -s 12 -c 6 -o gpurun_out/prof_mlp2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -8
