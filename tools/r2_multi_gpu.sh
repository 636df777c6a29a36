#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 500 --warmup 30 > gpurun_out/bench_n${N}_r2.json 2> gpurun_out/bench_n${N}_r2.err; tail -c 2500 gpurun_out/bench_n${N}_r2.json; grep -i "error\|Traceback" -A5 gpurun_out/bench_n${N}_r2.err | head -20
MRH_BENCH_INGEST=broadcast timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/bench_n${N}_bcast_r2.json 2> gpurun_out/bench_n${N}_bcast_r2.err; python - <<PY
import json
for f in ["bench_n${N}_r2", "bench_n${N}_bcast_r2"]:
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "us/step", round(d["e2e"]["ms_per_step"] * 1e3, 1))
    except Exception as e:
        print(f, "ERR", e)
PY
