#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_dw.py -q -x -m gpu > $O/pytest_dw.log 2>&1; echo "pytest dw rc=$?"; tail -5 $O/pytest_dw.log | cut -c1-300
timeout 300 python tools/gpu/gpu_probe_stash_io.py > $O/stash_io_probe.txt 2>&1; echo "probe rc=$?"; cat $O/stash_io_probe.txt | cut -c1-300
timeout 200 python tools/gpu/gpu_clk_tangent.py > $O/tangent_timeline.txt 2>&1; echo "rc=$?"; cat $O/tangent_timeline.txt | cut -c1-260
