#!/bin/bash
# One gpurun call: every GPU test file in its own process (a trapped kernel must not take the other files down),
# then a short C2 bench.  Logs land in gpurun_out/<tag>_*.
tag=${1:-run}
mkdir -p gpurun_out
for f in tests/test_packed_gpu.py tests/test_kernels_gpu.py tests/test_precise_gpu.py tests/test_parity_gpu.py tests/test_configs_gpu.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q -s --timeout 600 > gpurun_out/${tag}_${n}.log 2>&1
  echo "$n rc=$?"; tail -3 gpurun_out/${tag}_${n}.log
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","steps")}, d["e2e"]["value"], d["roofline"]["frac"], d.get("parity"))
    for k,v in d["kernels"].items(): print(k, v.get("avg_us"), v.get("launches_per_step"), v.get("frac_tensor"), v.get("frac_hbm"))
except Exception as e: print("bench parse failed", e)
PY
