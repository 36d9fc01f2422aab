#!/bin/bash
# Shortest gpurun call: GPU tests only (optionally a -k filter as $1).
mkdir -p gpurun_out
if [ -n "$1" ]; then
  timeout 600 python -m pytest tests -q -m gpu -k "$1" > gpurun_out/pytest_gpu.log 2>&1
else
  timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
fi
echo "pytest rc=$?"; grep -v "Warning\|warn\|^$\|Docs:" gpurun_out/pytest_gpu.log | tail -40
