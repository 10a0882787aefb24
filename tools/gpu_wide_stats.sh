#!/bin/bash
for k in ${1:-text sparse01 records}; do B200LZ4_LIB=build/libb200lz4_stats.so python tools/linked_probe.py --streams 128 --mib-per-stream 4 --kinds $k --stats 2>&1 | tail -2; done
