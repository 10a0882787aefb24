#!/bin/bash
# full visit: GPU suite, smoke, bench (+ reference arm), launch list.  usage: tools/gpu_full.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_$tag.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$tag.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"ceil",d["e2e"]["copy_ceiling"]["value"],"staged",d["e2e_staged"]["value"],"cpu",d["cpu_baseline"]["value"])
for k,v in d["extra"].items():
    if "value" in v: print(k, round(v["value"],1), v.get("e2e",{}).get("value"))
    else:
        print(k, {kk:(round(vv["value"],2), round(vv["roofline"]["frac"],4)) for kk,vv in v.items() if isinstance(vv,dict) and "value" in vv}, "ratio", round(v.get("ratio",0),2), v.get("round_trip_identical"))
PY
