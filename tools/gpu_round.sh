#!/bin/bash
# One GPU-box visit that regenerates the evidence under profiles/: parity suite, smoke, headline bench (+ reference arm),
# ncu launch list, ncu --set full of the four codec kernels (summarised on the box), per-kind probes, configs 3-5.
# usage: tools/gpu_round.sh <tag>      (outputs under gpurun_out/; copy what should be judged into profiles/)
# LIGHT=1: skip the ncu --set full captures and the kernel-resident probes (for visits that only changed the host pipeline)
# MEDIUM=1: as LIGHT, but keep the ncu --set full captures of compress_kernel and decompress_kernel
tag=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$tag.log
python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref rc=$?"
SHORT="python bench.py --steps 2 --warmup 3 --no-cpu --no-check --quick"
# (under ncu kernels run one at a time, so a kernel that waits for another kernel's flag would only time out: plain pipeline there)
export B200LZ4_NO_STREAMED=1
$SHORT > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv $SHORT > gpurun_out/ncu_launches_$tag.log 2>&1
echo "launch list rc=$?"
prof() {   # name, kernel regex, launches to skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/prof_${name}_$tag -f "$@" > gpurun_out/ncu_${name}_$tag.log 2>&1
  echo "ncu $name rc=$?"
  python tools/summarise_ncu.py gpurun_out/prof_${name}_$tag.ncu-rep gpurun_out/${tag}_${name}_ncu.txt --top 45 > /dev/null 2>&1
  rm -f gpurun_out/prof_${name}_$tag.ncu-rep
}
[ -n "$MEDIUM" ] && LIGHT=
if [ -z "$LIGHT" ]; then
prof compress_kernel 'compress_kernel$' 3 $SHORT
prof decompress_kernel 'decompress_kernel' 3 $SHORT
fi
[ -n "$MEDIUM" ] && LIGHT=1
if [ -z "$LIGHT" ]; then
prof decompress_kernel_wide_text decompress_kernel_wide 1 python tools/linked_probe.py --streams 128 --mib-per-stream 4 --kinds text
prof decompress_kernel_wide_mixed decompress_kernel_wide 1 python tools/linked_probe.py --streams 128 --mib-per-stream 4 --kinds mixed
prof compress_kernel_wide_mixed compress_kernel_wide 1 python tools/linked_probe.py --streams 128 --mib-per-stream 4 --kinds mixed
fi
unset B200LZ4_NO_STREAMED
python tools/streamed_probe.py --sweep auto,0,1000 2>&1 | grep -v "piece g\|group \|\[b200" > gpurun_out/${tag}_streamed_probe.txt
[ -n "$LIGHT" ] && { python tools/config_runs.py > gpurun_out/${tag}_configs_3_4_5.json 2> gpurun_out/configs_$tag.err; echo "configs rc=$?"; exit 0; }
python tools/kernel_probe.py --mib 1024 --kinds mixed,text,sparse01,records,random --blocks 65536,640000,4194304 --accels 1,400 > gpurun_out/${tag}_kernel_probe.txt 2>&1
for k in mixed text sparse01 records random zero; do python tools/linked_probe.py --streams 128 --mib-per-stream 8 --kinds $k 2>&1 | tail -1; done > gpurun_out/${tag}_linked_probe.txt
python tools/linked_probe.py --streams 1 --mib-per-stream 16 --kinds text,mixed 2>&1 | tail -2 >> gpurun_out/${tag}_linked_probe.txt
python tools/config_runs.py > gpurun_out/${tag}_configs_3_4_5.json 2> gpurun_out/configs_$tag.err; echo "configs rc=$?"
python tools/gather_bench.py > gpurun_out/${tag}_gather_bench.txt 2>&1
cat gpurun_out/${tag}_linked_probe.txt
