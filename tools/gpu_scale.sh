#!/bin/bash
# scaling evidence at N GPUs (gpurun --gpus N): BASELINE configs c2 / c3 / c5 through bench.py under torchrun
tag=${1:-scale}; N=${2:-8}; shift 2
mkdir -p gpurun_out
for cfg in "$@"; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 20 --warmup 5 --config $cfg > gpurun_out/${tag}_bench_${cfg}_n$N.json 2> gpurun_out/${tag}_bench_${cfg}_n$N.err
echo "bench $cfg N=$N rc=$?"; grep -v "^\[W\|^W0\|OMP_NUM\|\*\*\*\*" gpurun_out/${tag}_bench_${cfg}_n$N.err | tail -c 400
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_${cfg}_n$N.json").read().strip().splitlines()[-1])
    print("$cfg", {k:d[k] for k in ("value","ms_per_step","steps","n_gpus")}, "e2e", d["e2e"]["value"], "batch", d["config"]["global_batch"], d["clocks"])
except Exception as e: print("bench parse failed", e)
PY
done
