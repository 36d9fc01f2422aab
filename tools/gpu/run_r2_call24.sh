#!/bin/bash
# round 2, call 24: TMA-staged backward kernels + TMA training stash: full suite, bench lines, launch list, K1r ncu.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log | cut -c1-600
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
for wl in c2 c3; do timeout 300 python bench.py --workload $wl --graph --steps 30 --no-cpu-baseline --no-gpu-incumbent > $O/bench_$wl.json 2> $O/bench_$wl.err; echo "bench $wl rc=$?"; show $O/bench_$wl.json; done
timeout 300 python bench.py --rays 512 --graph --steps 30 --no-cpu-baseline --no-gpu-incumbent > $O/bench_train_512rays.json 2> $O/bench_train_512.err; echo "bench 512 rc=$?"; show $O/bench_train_512rays.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 160 --csv --log-file $O/launches_train.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-incumbent > $O/ncu_launches.log 2>&1; echo "ncu list rc=$?"
for mode in infer train; do
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:"mlp_rgrad" -s 3 -c 1 -o /tmp/prof_k1r_$mode python bench.py --mode $mode --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-incumbent > $O/ncu_k1r_$mode.log 2>&1; echo "ncu k1r $mode rc=$?"
  ncu -i /tmp/prof_k1r_$mode.ncu-rep --page raw --csv > $O/prof_k1r_${mode}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_k1r_$mode.ncu-rep --page source --csv 2>/dev/null | python tools/gpu/ncu_stalls.py > $O/prof_k1r_${mode}_stalls.txt
done
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"mlp_revt|weight_grad" -s 6 -c 2 -o /tmp/prof_bwd python bench.py --mode train --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-incumbent > $O/ncu_bwd.log 2>&1; echo "ncu bwd rc=$?"
ncu -i /tmp/prof_bwd.ncu-rep --page raw --csv > $O/prof_bwd_raw.csv 2>/dev/null
ncu -i /tmp/prof_bwd.ncu-rep --page source --csv 2>/dev/null | python tools/gpu/ncu_stalls.py > $O/prof_bwd_stalls.txt
