#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/gpu/gpu_clk_tangent.py > gpurun_out/tangent_timeline.txt 2>&1; echo "rc=$?"; cat gpurun_out/tangent_timeline.txt | cut -c1-260
