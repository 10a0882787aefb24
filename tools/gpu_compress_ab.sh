#!/bin/bash
# A/B of compressor builds: parity subset with the new build, then kernel probes with each library.
# usage: tools/gpu_compress_ab.sh "<lib1> <lib2>"
timeout 1500 python -m pytest -x -q -m gpu tests/test_gpu_fuzz.py tests/test_gpu_parity.py tests/test_gpu_modes.py -k "random_round_trips or echoing or split_over or generators or edge_sizes or acceleration_sweep or linked_state or forced_kernel or empty_and_tiny or quickcheck" > gpurun_out/pytest_cab.log 2>&1; echo "parity rc=$?"; tail -4 gpurun_out/pytest_cab.log
for lib in $1; do echo "== $lib"
  B200LZ4_LIB=$lib timeout 300 python tools/kernel_probe.py --mib 1024 --kinds mixed --blocks 640000 --accels 400,1 --reps 5 2>&1 | tail -2
  B200LZ4_LIB=$lib timeout 300 python tools/kernel_probe.py --mib 1024 --kinds text,sparse01,records --blocks 640000 --accels 1 --reps 3 2>&1 | tail -3
  B200LZ4_LIB=$lib timeout 300 python tools/kernel_probe.py --mib 1024 --kinds text --blocks 65536 --accels 1 --reps 3 2>&1 | tail -1
  B200LZ4_LIB=$lib timeout 300 python tools/linked_probe.py --streams 128 --mib-per-stream 4 --kinds mixed,text 2>&1 | tail -2
done
