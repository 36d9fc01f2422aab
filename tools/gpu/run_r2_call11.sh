#!/bin/bash
# round 2, call 11: rolled issuer default in K1r; rolled variants of K1 / tangent forward / reverse sweep / K1r
# reverse epilogue measured; full suite; bench lines; ncu of K1r (train) for the instruction-cache stalls.
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python tools/gpu/gpu_time_stages.py > $O/stages_time.txt 2>&1; echo "stages rc=$?"; cut -c1-330 $O/stages_time.txt
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log | cut -c1-600
show() { python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "ms", round(d["ms_per_step"], 3), "value", round(d["value"] / 1e6, 2), "M  e2e ms", round(d["e2e"]["ms_per_step"], 3),
          " frac", round(d["roofline"]["frac"], 4), "k_ms", round(d["roofline"]["ms_per_launch"], 3), "launches", d["gpu_launches"])
    if d.get("graphed"): print("   graphed:", json.dumps(d.get("graphed"))[:260])
except Exception as e:
    print(f, "unreadable", e)
PY
}
timeout 400 python bench.py --graph --no-cpu-baseline --no-gpu-incumbent > $O/bench_train_fp32.json 2> $O/bench_train.err; echo "bench train rc=$?"; show $O/bench_train_fp32.json; tail -2 $O/bench_train.err | cut -c1-300
timeout 300 python bench.py --mode infer --graph --no-cpu-baseline --no-gpu-incumbent > $O/bench_infer_fp32.json 2> $O/bench_infer.err; echo "bench infer rc=$?"; show $O/bench_infer_fp32.json
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"mlp_rgrad" -s 3 -c 1 -o /tmp/prof_k1r_train python bench.py --mode train --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-incumbent > $O/ncu_k1r_train.log 2>&1; echo "ncu k1r train rc=$?"
ncu -i /tmp/prof_k1r_train.ncu-rep --page raw --csv > $O/prof_k1r_train_raw.csv 2>/dev/null
ncu -i /tmp/prof_k1r_train.ncu-rep --page source --csv 2>/dev/null | python tools/gpu/ncu_stalls.py > $O/prof_k1r_train_stalls.txt
head -12 $O/prof_k1r_train_stalls.txt
