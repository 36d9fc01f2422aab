#!/bin/bash
# round 2, final validation: what the driver runs at round end (GPU tests, smoke, both bench arms with defaults),
# the other BASELINE configs, and (with FULL=1) the ncu captures of K1r + the launch list of a train step.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log | cut -c1-300
( time timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench_default.time; echo "bench default rc=$?"; cut -c1-260 $O/bench_default.json; grep real $O/bench_default.time
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err ) 2> $O/bench_reference.time; echo "bench reference rc=$?"; cut -c1-200 $O/bench_reference.json; grep real $O/bench_reference.time
timeout 300 python bench.py --mode infer --graph --no-cpu-baseline --no-gpu-incumbent > $O/bench_infer_fp32.json 2> $O/bench_infer.err; echo "bench infer rc=$?"; cut -c1-200 $O/bench_infer_fp32.json
for wl in c2 c3; do timeout 300 python bench.py --workload $wl --graph --steps 30 --no-cpu-baseline --no-gpu-incumbent > $O/bench_$wl.json 2> $O/bench_$wl.err; echo "bench $wl rc=$?"; cut -c1-200 $O/bench_$wl.json; done
timeout 300 python bench.py --rays 512 --graph --steps 30 --no-cpu-baseline --no-gpu-incumbent > $O/bench_train_512rays.json 2> $O/bench_train_512.err; echo "bench 512 rc=$?"; cut -c1-200 $O/bench_train_512rays.json
timeout 300 python bench.py --workload c5 --mode infer --steps 5 --no-cpu-baseline --no-gpu-incumbent > $O/bench_c5.json 2> $O/bench_c5.err; echo "bench c5 rc=$?"; cut -c1-200 $O/bench_c5.json
if [ -n "$FULL" ]; then
for mode in infer train; do
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:"mlp_rgrad" -s 3 -c 1 -o /tmp/prof_k1r_$mode python bench.py --mode $mode --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-incumbent > $O/ncu_k1r_$mode.log 2>&1; echo "ncu k1r $mode rc=$?"
  ncu -i /tmp/prof_k1r_$mode.ncu-rep --page raw --csv > $O/prof_k1r_${mode}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_k1r_$mode.ncu-rep --page source --csv 2>/dev/null | python tools/gpu/ncu_stalls.py > $O/prof_k1r_${mode}_stalls.txt
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 160 --csv --log-file $O/launches_train.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-incumbent > $O/ncu_launches.log 2>&1; echo "ncu list rc=$?"
fi
