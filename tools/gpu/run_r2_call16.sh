#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu/gpu_probe_stash_io.py > gpurun_out/stash_io_probe.txt 2>&1; echo "rc=$?"; cat gpurun_out/stash_io_probe.txt | cut -c1-300
