#!/bin/bash
# short multi-GPU visit: the bench as the driver launches it (first: it is what the round-end scaling run executes), then the mctx tests
# usage: tools/gpu_scale_quick.sh <N> <tag>
N=$1; tag=$2
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout ${BENCH_TIMEOUT:-330} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_${tag}.err | cut -c1-300
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/bench_${tag}.json") if l.startswith("{")][-1]
print("N",d["n_gpus"],"value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"plain",round(d["e2e"]["plain_pipeline"]["value"],1),"ceil",round(d["e2e"]["copy_ceiling"]["value"],1),"frac",round(d["e2e"]["frac_of_copy_ceiling"],2),"staged",d["e2e_staged"] and round(d["e2e_staged"]["value"],1))
for k in ("mctx_one_process","config3_strong","d+640000"):
    print(k, json.dumps(d["extra"].get(k))[:700])
PY
[ -n "$NO_TESTS" ] || timeout 200 python -m pytest -x -q -m gpu tests/test_gpu_multi.py > gpurun_out/pytest_multi_${tag}.log 2>&1; echo "multi tests rc=$?"; tail -3 gpurun_out/pytest_multi_${tag}.log
