#!/bin/bash
# quick check of selected tests under both decoder kernels.  usage: tools/gpu_quick.sh "<-k expression>"
for w in "" 0 1; do
  echo "== B200LZ4_DWIDE=$w"
  if [ -z "$w" ]; then timeout 900 python -m pytest -x -q -m gpu tests/test_gpu_fuzz.py tests/test_gpu_parity.py -k "$1" 2>&1 | tail -6
  else B200LZ4_DWIDE=$w timeout 900 python -m pytest -x -q -m gpu tests/test_gpu_fuzz.py tests/test_gpu_parity.py -k "$1" 2>&1 | tail -6; fi
done
