#!/bin/bash
# round 2, call 1: the new defaults (K1r + shared-forward backward) -- tests, bench lines, K1r timeline, ncu.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 900 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
timeout 300 python bench.py > $O/bench_train_fp32.json 2> $O/bench_train.err; echo "bench train rc=$?"; cut -c1-400 $O/bench_train_fp32.json
timeout 200 python bench.py --mode infer --no-cpu-baseline > $O/bench_infer_fp32.json 2> $O/bench_infer.err; echo "bench infer rc=$?"; cut -c1-300 $O/bench_infer_fp32.json
timeout 120 python tools/gpu/gpu_clk_rgrad.py > $O/k1r_clk.txt 2>&1; echo "clk rc=$?"
timeout 200 python tools/gpu/gpu_time_rgrad.py > $O/k1r_time.txt 2>&1; echo "time rc=$?"; cat $O/k1r_time.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file $O/launches_train.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"mlp_rgrad" -s 3 -c 1 -o /tmp/prof_k1r python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_k1r.log 2>&1; echo "ncu k1r rc=$?"
ncu -i /tmp/prof_k1r.ncu-rep --page raw --csv > $O/prof_k1r_raw.csv 2>/dev/null
ncu -i /tmp/prof_k1r.ncu-rep --page source --csv 2>/dev/null | python tools/gpu/ncu_stalls.py > $O/prof_k1r_stalls.txt
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"mlp_kernel<1, 3|mlp_rev|dual_top|db_partial" -s 12 -c 4 -o /tmp/prof_bwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_bwd.log 2>&1; echo "ncu bwd rc=$?"
ncu -i /tmp/prof_bwd.ncu-rep --page raw --csv > $O/prof_bwd_raw.csv 2>/dev/null
ncu -i /tmp/prof_bwd.ncu-rep --page source --csv 2>/dev/null | python tools/gpu/ncu_stalls.py > $O/prof_bwd_stalls.txt
ls -la $O | tail -30; du -sh $O
