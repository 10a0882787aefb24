#!/bin/bash
# wide-decoder check: forced-wide parity subset, then linked throughput per data kind.  usage: tools/gpu_wide_check.sh <tag> [full]
# (tight timeouts: a pipeline bug in the wide kernel shows up as a hang, and GPU minutes are scarce)
tag=${1:-x}
mkdir -p gpurun_out
timeout 60 python tools/linked_probe.py --streams 4 --mib-per-stream 1 --kinds text 2>&1 | tail -1 || { echo "SMOKE HANG/FAIL"; exit 1; }
B200LZ4_DWIDE=1 timeout 240 python -m pytest -x -q -m gpu tests/test_gpu_fuzz.py tests/test_gpu_parity.py tests/test_gpu_configs.py -k "handbuilt or corrupted or random_round_trips or echoing or split_over or generators or edge_sizes or empty_and_tiny or block_max or large_blocks or linked_state or malformed or fragmented or config4 or config3" > gpurun_out/pytest_wide_$tag.log 2>&1; rc=$?; echo "forced-wide rc=$rc"; tail -4 gpurun_out/pytest_wide_$tag.log
[ $rc -ne 0 ] && exit 1
for k in mixed text sparse01 records random zero; do timeout 60 python tools/linked_probe.py --streams 128 --mib-per-stream 4 --kinds $k 2>&1 | tail -1; done | tee gpurun_out/linked_kinds_$tag.log
timeout 60 python tools/linked_probe.py --streams 1 --mib-per-stream 16 --kinds text,mixed 2>&1 | tail -2 | tee -a gpurun_out/linked_kinds_$tag.log
if [ "$2" = "full" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$tag.log
fi
