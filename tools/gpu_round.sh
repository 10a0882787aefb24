#!/bin/bash
# One GPU-box visit: parity suite, headline bench, ncu launch list, ncu --set full of the two codec kernels.
# usage: tools/gpu_round.sh <tag>      (outputs under gpurun_out/, summarised into profiles/ by tools/summarise_ncu.py)
tag=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref rc=$?"
SHORT="python bench.py --steps 2 --warmup 3 --no-cpu --no-check"
$SHORT > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $SHORT > gpurun_out/ncu_launches_$tag.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:compress_kernel -s 3 -c 1 -o gpurun_out/prof_compress_$tag -f $SHORT > gpurun_out/ncu_c_$tag.log 2>&1
echo "ncu compress rc=$?"
ncu --set full --clock-control none --import-source on -k regex:decompress_kernel -s 3 -c 1 -o gpurun_out/prof_decompress_$tag -f $SHORT > gpurun_out/ncu_d_$tag.log 2>&1
echo "ncu decompress rc=$?"
python tools/kernel_probe.py --mib 512 --kinds mixed,text --blocks 65536,640000,4194304 --accels 1,400 > gpurun_out/probe_$tag.log 2>&1
tail -20 gpurun_out/probe_$tag.log
cat gpurun_out/bench_$tag.json
