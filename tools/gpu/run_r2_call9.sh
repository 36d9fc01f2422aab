#!/bin/bash
# round 2, call 9: stash gating fixed (grad mode sampled outside Function.forward), rolled-issuer K1r A/B,
# CUDA-graph iteration, full suite, bench lines, ncu of K1r in train mode (with the stash).
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/gpu/gpu_time_rgrad.py > $O/k1r_variants_time.txt 2>&1; echo "rgrad rc=$?"; grep -v "k1_dot\|persisting-L2 window\|N-split" $O/k1r_variants_time.txt | cut -c1-220
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_gpu.log | cut -c1-300
timeout 400 python bench.py --graph --no-cpu-baseline --no-gpu-incumbent > $O/bench_train_fp32.json 2> $O/bench_train.err; echo "bench train rc=$?"; cut -c1-200 $O/bench_train_fp32.json; tail -3 $O/bench_train.err
python - <<'PY'
import json
for f in ("bench_train_fp32",):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "graphed:", d.get("graphed"), "e2e:", d["e2e"]["ms_per_step"], "roof:", d["roofline"]["frac"], d["roofline"]["ms_per_launch"], "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 300 python bench.py --mode infer --graph --no-cpu-baseline --no-gpu-incumbent > $O/bench_infer_fp32.json 2> $O/bench_infer.err; echo "bench infer rc=$?"; cut -c1-200 $O/bench_infer_fp32.json; grep -o '"graphed": {[^}]*}' $O/bench_infer_fp32.json | cut -c1-400
timeout 300 python bench.py --rays 512 --graph --steps 30 --no-cpu-baseline --no-gpu-incumbent > $O/bench_train_512rays.json 2> $O/bench_train_512.err; echo "bench 512 rc=$?"; cut -c1-200 $O/bench_train_512rays.json; grep -o '"graphed": {[^}]*}' $O/bench_train_512rays.json | cut -c1-400; tail -2 $O/bench_train_512.err
timeout 300 python bench.py --workload c2 --graph --steps 30 --no-cpu-baseline --no-gpu-incumbent > $O/bench_c2.json 2> $O/bench_c2.err; echo "bench c2 rc=$?"; cut -c1-200 $O/bench_c2.json; grep -o '"graphed": {[^}]*}' $O/bench_c2.json | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 160 --csv --log-file $O/launches_train.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-incumbent > $O/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"mlp_rgrad" -s 3 -c 1 -o /tmp/prof_k1r_train python bench.py --mode train --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-incumbent > $O/ncu_k1r_train.log 2>&1; echo "ncu k1r train rc=$?"
ncu -i /tmp/prof_k1r_train.ncu-rep --page raw --csv > $O/prof_k1r_train_raw.csv 2>/dev/null
ncu -i /tmp/prof_k1r_train.ncu-rep --page source --csv 2>/dev/null | python tools/gpu/ncu_stalls.py > $O/prof_k1r_train_stalls.txt
ls -la $O | tail -8
