#!/bin/bash
# round 2, visit B: bench retry + ncu of the decoder on text and of the wide compressor on linked streams
tag=r2b
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$tag.err
D="python tools/kernel_probe.py --mib 1024 --kinds text --blocks 640000 --accels 1 --reps 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decompress_kernel -s 1 -c 1 -o gpurun_out/prof_dtext_$tag -f $D > gpurun_out/ncu_dtext_$tag.log 2>&1; echo "ncu dtext rc=$?"
python tools/summarise_ncu.py gpurun_out/prof_dtext_$tag.ncu-rep gpurun_out/dtext_$tag.txt --top 70 > /dev/null 2>&1
W="python tools/linked_probe.py --streams 128 --mib-per-stream 4 --kinds mixed"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:compress_kernel_wide -s 1 -c 1 -o gpurun_out/prof_cwide_$tag -f $W > gpurun_out/ncu_cwide_$tag.log 2>&1; echo "ncu cwide rc=$?"
python tools/summarise_ncu.py gpurun_out/prof_cwide_$tag.ncu-rep gpurun_out/cwide_$tag.txt --top 90 > /dev/null 2>&1
rm -f gpurun_out/prof_dtext_$tag.ncu-rep gpurun_out/prof_cwide_$tag.ncu-rep
cat gpurun_out/bench_$tag.json | cut -c1-3000
