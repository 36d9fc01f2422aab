#!/bin/bash
# First GPU call of the next round: bring up K1r (mlp_rg.cu), which was written after round 1's GPU
# budget was spent.  Parity first (bounded by pytest-timeout; a protocol bug traps after ~4 s per
# mbarrier wait instead of hanging), then timing against K1g, bench lines in both grad modes, ncu.
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/gpu/gpu_debug_rgrad.py > $O/k1r_debug.txt 2>&1; echo "k1r debug rc=$?"; tail -60 $O/k1r_debug.txt
EMAP_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_rgrad.py -x -q -k "not k1_dot and not shared" > $O/k1r_pytest.log 2>&1; echo "k1r pytest rc=$?"; tail -15 $O/k1r_pytest.log
EMAP_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_rgrad.py -x -q -k "shared" > $O/shared_pytest.log 2>&1; echo "shared-backward pytest rc=$?"; tail -15 $O/shared_pytest.log
EMAP_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_rgrad.py -x -q -k "k1_dot" > $O/k1dot_pytest.log 2>&1; echo "k1_dot pytest rc=$?"; tail -5 $O/k1dot_pytest.log
EMAP_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_rev2.py -x -q > $O/rev2_pytest.log 2>&1; echo "rev2 pytest rc=$?"; tail -5 $O/rev2_pytest.log
timeout 200 python tools/gpu/gpu_time_rev2.py > $O/rev2_time.txt 2>&1; echo "rev2 time rc=$?"; cat $O/rev2_time.txt
timeout 600 python -m pytest tests -q -m gpu --deselect tests/test_gpu_zz_experimental.py > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 200 python tools/gpu/gpu_time_rgrad.py > $O/k1r_time.txt 2>&1; echo "time rc=$?"; cat $O/k1r_time.txt
timeout 120 python tools/gpu/gpu_clk_rgrad.py > $O/k1r_clk.txt 2>&1; echo "clk rc=$?"
for gm in forward reverse; do
  timeout 200 python bench.py --mode infer --no-cpu-baseline --no-experimental --grad-mode $gm > $O/bench_infer_fp32_$gm.json 2> $O/bench_infer_$gm.err; echo "bench infer $gm rc=$?"; cut -c1-200 $O/bench_infer_fp32_$gm.json
  timeout 300 python bench.py --no-cpu-baseline --no-experimental --grad-mode $gm > $O/bench_train_fp32_$gm.json 2> $O/bench_train_$gm.err; echo "bench train $gm rc=$?"; cut -c1-200 $O/bench_train_fp32_$gm.json
done
timeout 300 python bench.py --no-cpu-baseline --no-experimental --grad-mode reverse --bwd-stash shared > $O/bench_train_fp32_reverse_shared.json 2> $O/bench_train_shared.err; echo "bench train reverse+shared rc=$?"; cut -c1-200 $O/bench_train_fp32_reverse_shared.json
timeout 300 python bench.py --no-cpu-baseline --no-experimental --all-optins > $O/bench_train_fp32_all_optins.json 2> $O/bench_train_all.err; echo "bench train all opt-ins rc=$?"; cut -c1-200 $O/bench_train_fp32_all_optins.json
timeout 200 python bench.py --mode infer --no-cpu-baseline --no-experimental --all-optins > $O/bench_infer_fp32_all_optins.json 2>/dev/null; echo "bench infer all opt-ins rc=$?"; cut -c1-200 $O/bench_infer_fp32_all_optins.json
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"mlp_rgrad" -s 3 -c 1 -o /tmp/prof_k1r python bench.py --mode infer --steps 1 --warmup 3 --no-cpu-baseline --no-experimental --grad-mode reverse > $O/ncu_k1r.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/prof_k1r.ncu-rep --page raw --csv > $O/prof_k1r_raw.csv 2>/dev/null
ncu -i /tmp/prof_k1r.ncu-rep --page source --csv 2>/dev/null | python tools/gpu/ncu_stalls.py > $O/prof_k1r_stalls.txt
ls -la $O | tail; du -sh $O
