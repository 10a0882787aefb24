#!/bin/bash
# round 2, visit A: new tests + new bench + decoder diagnostics
tag=r2a
mkdir -p gpurun_out
nvidia-smi -L | head -3; nproc; free -g | head -2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$tag.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$tag.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref rc=$?"
echo "--- decoder diagnostics: normal vs copier-skips-copies (parser-bound rate)"
timeout 300 python tools/linked_probe.py --streams 128 --mib-per-stream 8 --kinds mixed,text > gpurun_out/linked_$tag.log 2>&1; cat gpurun_out/linked_$tag.log
B200LZ4_DECODE_DEBUG=1 timeout 300 python tools/linked_probe.py --streams 128 --mib-per-stream 8 --kinds mixed,text > gpurun_out/linked_nocopy_$tag.log 2>&1; cat gpurun_out/linked_nocopy_$tag.log
timeout 300 python tools/kernel_probe.py --mib 1024 --kinds mixed,text --blocks 640000,4194304 --accels 1 > gpurun_out/probe_$tag.log 2>&1; cat gpurun_out/probe_$tag.log
B200LZ4_DECODE_DEBUG=1 timeout 300 python tools/kernel_probe.py --mib 1024 --kinds mixed,text --blocks 640000,4194304 --accels 1 > gpurun_out/probe_nocopy_$tag.log 2>&1; cat gpurun_out/probe_nocopy_$tag.log
cat gpurun_out/bench_$tag.json | cut -c1-6000
