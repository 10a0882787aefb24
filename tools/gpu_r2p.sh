#!/bin/bash
tag=r2p
bash tools/gpu_wide_check.sh $tag
B200LZ4_DWIDE=0 timeout 1200 python -m pytest -x -q -m gpu tests/test_gpu_fuzz.py tests/test_gpu_parity.py -k "handbuilt or corrupted or random_round_trips or echoing or split_over or generators or edge_sizes or empty_and_tiny or block_max or large_blocks or linked_state or malformed or fragmented" > gpurun_out/pytest_narrow_$tag.log 2>&1; echo "forced-narrow rc=$?"; tail -3 gpurun_out/pytest_narrow_$tag.log
timeout 300 python tools/kernel_probe.py --mib 1024 --kinds mixed,text --blocks 65536,640000,4194304 --accels 1 2>&1 | tail -7
B200LZ4_DECODE_DEBUG=1 timeout 300 python tools/kernel_probe.py --mib 1024 --kinds text --blocks 640000 --accels 1 2>&1 | tail -1
