#!/bin/bash
# GPU visit for the streamed host calls: schedule / copy-mode matrix on config 2.
mkdir -p gpurun_out
timeout 240 python tools/streamed_probe.py --matrix > gpurun_out/streamed_matrix.log 2>&1; echo "matrix rc=$?"; grep -v "piece g\|group " gpurun_out/streamed_matrix.log | tail -60
