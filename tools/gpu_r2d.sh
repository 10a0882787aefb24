#!/bin/bash
tag=r2d
mkdir -p gpurun_out
W="python tools/linked_probe.py --streams 128 --mib-per-stream 4 --kinds text"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decompress_kernel_wide -s 1 -c 1 -o gpurun_out/prof_dwide_$tag -f $W > gpurun_out/ncu_dwide_$tag.log 2>&1; echo "ncu dwide rc=$?"
python tools/summarise_ncu.py gpurun_out/prof_dwide_$tag.ncu-rep gpurun_out/dwide_text_$tag.txt --top 70 > /dev/null 2>&1
W="python tools/linked_probe.py --streams 128 --mib-per-stream 4 --kinds mixed"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decompress_kernel_wide -s 1 -c 1 -o gpurun_out/prof_dwidem_$tag -f $W > gpurun_out/ncu_dwidem_$tag.log 2>&1; echo "ncu dwide mixed rc=$?"
python tools/summarise_ncu.py gpurun_out/prof_dwidem_$tag.ncu-rep gpurun_out/dwide_mixed_$tag.txt --top 50 > /dev/null 2>&1
rm -f gpurun_out/prof_dwide*_$tag.ncu-rep
for k in text random sparse01 records; do timeout 200 python tools/linked_probe.py --streams 128 --mib-per-stream 4 --kinds $k 2>&1 | tail -1; done
