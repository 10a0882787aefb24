#!/bin/bash
# A/B of wide-decoder builds: tools/gpu_wide_ab.sh "<lib1> <lib2> ..." "<kinds>"
for lib in $1; do echo "== $lib"; for k in ${2:-text sparse01 records mixed}; do B200LZ4_LIB=$lib python tools/linked_probe.py --streams 128 --mib-per-stream 4 --kinds $k --stats 2>&1 | tail -2 | cut -c1-900; done; done
