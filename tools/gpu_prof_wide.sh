#!/bin/bash
# ncu source-level profile of decompress_kernel_wide on one data kind.  usage: tools/gpu_prof_wide.sh <tag> <kind>
tag=$1; kind=${2:-text}
W="python tools/linked_probe.py --streams 128 --mib-per-stream 4 --kinds $kind"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decompress_kernel_wide -s 1 -c 1 -o gpurun_out/prof_dwide_$tag -f $W > gpurun_out/ncu_dwide_$tag.log 2>&1; echo "ncu rc=$?"
python tools/summarise_ncu.py gpurun_out/prof_dwide_$tag.ncu-rep gpurun_out/dwide_${kind}_$tag.txt --top 60 > /dev/null 2>&1
rm -f gpurun_out/prof_dwide_$tag.ncu-rep
