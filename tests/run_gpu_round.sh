#!/bin/bash
# One gpurun call: tests, smoke, bench lines, ncu launch list + one full capture of the top kernel.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --mode infer > gpurun_out/bench_infer_fp32.json 2> gpurun_out/bench_infer_fp32.err; echo "bench rc=$?"; cat gpurun_out/bench_infer_fp32.json
timeout 300 python bench.py --steps 10 --warmup 3 --mode infer --precision fp16 --no-cpu-baseline > gpurun_out/bench_infer_fp16.json 2> gpurun_out/bench_infer_fp16.err; cat gpurun_out/bench_infer_fp16.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --mode infer --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mlp_kernel -s 8 -c 2 -o gpurun_out/prof_mlp python bench.py --steps 1 --warmup 3 --mode infer --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -12
