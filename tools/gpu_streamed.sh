#!/bin/bash
# GPU visit for the streamed host calls: schedule matrix on config 2, decode output modes, parity.
mkdir -p gpurun_out
timeout 240 python tools/streamed_probe.py --matrix > gpurun_out/streamed_matrix.log 2>&1; echo "matrix rc=$?"; grep -v "piece g\|group " gpurun_out/streamed_matrix.log | tail -40
timeout 600 python -m pytest -x -q -m gpu tests/test_gpu_streamed.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_configs.py > gpurun_out/pytest_subset.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_subset.log
