#!/bin/bash
# round 2, visit C: first run of the wide decoder
tag=r2c
mkdir -p gpurun_out
echo "--- wide decoder, linked probe"
timeout 300 python tools/linked_probe.py --streams 128 --mib-per-stream 8 --kinds mixed,text > gpurun_out/linked_$tag.log 2>&1; cat gpurun_out/linked_$tag.log | tail -4
timeout 120 python tools/linked_probe.py --streams 1 --mib-per-stream 16 --kinds text > gpurun_out/linked1_$tag.log 2>&1; cat gpurun_out/linked1_$tag.log | tail -2
echo "--- forced wide on the fuzz subset"
B200LZ4_DWIDE=1 timeout 1200 python -m pytest -x -q -m gpu tests/test_gpu_fuzz.py tests/test_gpu_parity.py -k "handbuilt or corrupted or random_round_trips or echoing or split_over or generators or edge_sizes or empty_and_tiny or block_max or large_blocks or linked_state or malformed or fragmented" > gpurun_out/pytest_wide_$tag.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_wide_$tag.log
echo "--- whole suite"
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$tag.log
timeout 300 python tools/kernel_probe.py --mib 1024 --kinds mixed,text --blocks 4194304 --accels 1 > gpurun_out/probe_$tag.log 2>&1; cat gpurun_out/probe_$tag.log
