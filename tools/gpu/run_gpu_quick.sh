#!/bin/bash
# Short gpurun call: smoke + GPU tests + kernel timing decomposition (no ncu, no bench).
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 600 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_gpu.log
timeout 120 python tools/gpu/gpu_time_mlp.py > $O/time_mlp.txt 2>&1; echo "time_mlp rc=$?"; cat $O/time_mlp.txt
timeout 120 python tools/gpu/gpu_time_bwd.py > $O/time_bwd.txt 2>&1; echo "time_bwd rc=$?"; cat $O/time_bwd.txt
timeout 120 python tools/gpu/gpu_time_bwd2.py > $O/time_bwd2.txt 2>&1; echo "time_bwd2 rc=$?"; cat $O/time_bwd2.txt
timeout 120 python tools/gpu/gpu_clk_mlp.py > $O/clk_mlp.txt 2>&1; echo "clk rc=$?"
