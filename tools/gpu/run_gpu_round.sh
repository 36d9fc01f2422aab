#!/bin/bash
# One gpurun call: smoke, GPU tests, bench lines (both arms), kernel timing decomposition, ncu launch
# list + full captures of the top kernels exported to CSV/text on the box (the .ncu-rep files are too
# big to travel).  Everything lands in gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 600 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python bench.py > $O/bench_train_fp32.json 2> $O/bench_train.err; echo "bench train rc=$?"; cut -c1-300 $O/bench_train_fp32.json
timeout 200 python bench.py --mode infer --no-cpu-baseline > $O/bench_infer_fp32.json 2> $O/bench_infer.err; echo "bench infer rc=$?"; cut -c1-200 $O/bench_infer_fp32.json
timeout 200 python bench.py --mode infer --precision fp16 --no-cpu-baseline > $O/bench_infer_fp16.json 2>/dev/null
timeout 200 python bench.py --precision fp16 --no-cpu-baseline > $O/bench_train_fp16.json 2>/dev/null
for wl in c2 c3 c5; do timeout 300 python bench.py --workload $wl --no-cpu-baseline --steps 10 > $O/bench_$wl.json 2> $O/bench_$wl.err; echo "bench $wl rc=$?"; cut -c1-200 $O/bench_$wl.json; done
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2>/dev/null; echo "ref rc=$?"; cut -c1-200 $O/bench_reference.json
timeout 120 python tools/gpu/gpu_time_mlp.py > $O/time_mlp.txt 2>&1; echo "time_mlp rc=$?"
timeout 120 python tools/gpu/gpu_time_bwd.py > $O/time_bwd.txt 2>&1; echo "time_bwd rc=$?"
timeout 120 python tools/gpu/gpu_time_bwd2.py > $O/time_bwd2.txt 2>&1; echo "time_bwd2 rc=$?"
timeout 120 python tools/gpu/gpu_clk_mlp.py > $O/clk_mlp.txt 2>&1; echo "clk rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 160 --csv --log-file $O/launches_train.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"mlp_kernel|mlp_rev" -s 21 -c 7 -o /tmp/prof_mlp python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/prof_mlp.ncu-rep --page raw --csv > $O/prof_mlp_raw.csv 2>/dev/null
ncu -i /tmp/prof_mlp.ncu-rep --page source --csv 2>/dev/null | python tools/gpu/ncu_stalls.py > $O/prof_mlp_stalls.txt
ls -la $O | tail -20; du -sh $O
