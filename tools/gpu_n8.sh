#!/bin/bash
# 8-GPU round: the stated multi-GPU configurations through bench.py
mkdir -p gpurun_out
for c in c2 c3 c5; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 20 --warmup 5 --config $c > gpurun_out/n8_bench_$c.json 2> gpurun_out/n8_bench_$c.err; echo "$c rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/n8_bench_$c.json") if l.startswith("{")][-1])
    print("$c", {k:d[k] for k in ("value","n_gpus","ms_per_step","steps")}, "e2e", d["e2e"]["value"], d["clocks"]["sm_mhz"])
except Exception as e: print("parse failed", e)
PY
done
