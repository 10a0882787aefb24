#!/bin/bash
tag=r2q
timeout 900 python -m pytest -x -q -m gpu tests/test_gpu_frame.py tests/test_gpu_parity.py -k "frame or replayed or c_example or xxh32 or stock" > gpurun_out/pytest_new_$tag.log 2>&1; echo "new tests rc=$?"; tail -15 gpurun_out/pytest_new_$tag.log
timeout 300 python tools/gather_bench.py 2>&1 | tail -16
for w in 0 1; do B200LZ4_DWIDE=$w timeout 200 python tools/kernel_probe.py --mib 512 --kinds mixed,text --blocks 4194304 --accels 1 2>&1 | tail -2; done
