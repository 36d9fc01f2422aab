#!/bin/bash
# round 2, call 10: fold tables from pinned immortal memory (CUDA-graph safe), graph tests in child processes;
# full suite, graphed bench lines, K1r+stash variants.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_gpu.log | cut -c1-600
timeout 300 python tools/gpu/gpu_time_stages.py > $O/stages_time.txt 2>&1; echo "stages rc=$?"; cut -c1-300 $O/stages_time.txt
show() { python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "ms", round(d["ms_per_step"], 3), "value", round(d["value"] / 1e6, 2), "M  e2e ms", round(d["e2e"]["ms_per_step"], 3),
          " frac", round(d["roofline"]["frac"], 4), "k_ms", round(d["roofline"]["ms_per_launch"], 3), "launches", d["gpu_launches"])
    print("   graphed:", json.dumps(d.get("graphed"))[:500])
except Exception as e:
    print(f, "unreadable", e)
PY
}
timeout 400 python bench.py --graph --no-cpu-baseline --no-gpu-incumbent > $O/bench_train_fp32.json 2> $O/bench_train.err; echo "bench train rc=$?"; show $O/bench_train_fp32.json; tail -2 $O/bench_train.err | cut -c1-300
timeout 300 python bench.py --mode infer --graph --no-cpu-baseline --no-gpu-incumbent > $O/bench_infer_fp32.json 2> $O/bench_infer.err; echo "bench infer rc=$?"; show $O/bench_infer_fp32.json
timeout 300 python bench.py --rays 512 --graph --steps 30 --no-cpu-baseline --no-gpu-incumbent > $O/bench_train_512rays.json 2> $O/bench_train_512.err; echo "bench 512 rc=$?"; show $O/bench_train_512rays.json
timeout 300 python bench.py --workload c2 --graph --steps 30 --no-cpu-baseline --no-gpu-incumbent > $O/bench_c2.json 2> $O/bench_c2.err; echo "bench c2 rc=$?"; show $O/bench_c2.json
timeout 300 python bench.py --workload c3 --graph --steps 30 --no-cpu-baseline --no-gpu-incumbent > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench c3 rc=$?"; show $O/bench_c3.json
ls -la $O | tail -5
