#!/bin/bash
for lib in $1; do echo "== $lib"
  B200LZ4_LIB=$lib timeout 300 python tools/kernel_probe.py --mib 1024 --kinds mixed --blocks 640000 --accels 400,1 --reps 5 2>&1 | tail -2
  B200LZ4_LIB=$lib timeout 300 python tools/kernel_probe.py --mib 1024 --kinds text,sparse01,records --blocks 65536 --accels 1 --reps 3 2>&1 | tail -3
  B200LZ4_LIB=$lib timeout 300 python tools/kernel_probe.py --mib 4096 --kinds mixed --blocks 640000 --accels 1 --reps 2 2>&1 | tail -1
done
