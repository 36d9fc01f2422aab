#!/bin/bash
# round 2, call 13: ncu of the tangent forward (5th mlp_kernel launch of a train step), C5 at N=1, other configs.
mkdir -p gpurun_out
O=gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'^mlp_kernel$' -s 14 -c 1 -o /tmp/prof_tangent python bench.py --mode train --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-incumbent > $O/ncu_tangent.log 2>&1; echo "ncu tangent rc=$?"
ncu -i /tmp/prof_tangent.ncu-rep --page raw --csv > $O/prof_tangent_raw.csv 2>/dev/null
ncu -i /tmp/prof_tangent.ncu-rep --page source --csv 2>/dev/null | python tools/gpu/ncu_stalls.py > $O/prof_tangent_stalls.txt
head -40 $O/prof_tangent_stalls.txt | cut -c1-230
timeout 300 python bench.py --workload c5 --mode infer --steps 5 --no-cpu-baseline --no-gpu-incumbent > $O/bench_c5.json 2> $O/bench_c5.err; echo "bench c5 rc=$?"; cut -c1-200 $O/bench_c5.json
timeout 300 python bench.py --mode infer --no-cpu-baseline --no-gpu-incumbent > $O/bench_infer_fp32.json 2> $O/bench_infer.err; echo "bench infer rc=$?"; cut -c1-200 $O/bench_infer_fp32.json
