#!/bin/bash
# multi-GPU bench of configs[3] on N GPUs of one box: tools/r2_multi_gpu.sh N   (under gpurun --gpus N)
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 500 --warmup 30 --no-cpu-baseline > gpurun_out/bench_n${N}_r2.json 2> gpurun_out/bench_n${N}_r2.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_n${N}_r2.json") if l.startswith("{")][-1])
    print("N=${N} value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "us/step", round(d["e2e"]["ms_per_step"] * 1e3, 1), "replica", d.get("replica_streams", {}).get("value"), "mesh", d.get("sharded_mesh", {}).get("ms"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/bench_n${N}_r2.err").read()[-3000:])
PY
