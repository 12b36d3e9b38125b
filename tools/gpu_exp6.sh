#!/bin/bash
tag=${1:-exp6}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -x -k "pool_match or single_head or tie" > gpurun_out/${tag}_kernels.log 2>&1; echo "kernels rc=$?"; tail -3 gpurun_out/${tag}_kernels.log
timeout 600 python -m pytest tests/test_configs_gpu.py -m gpu -q --timeout 600 -x -k "c4" > gpurun_out/${tag}_c4.log 2>&1; echo "c4 rc=$?"; tail -3 gpurun_out/${tag}_c4.log
timeout 300 python tools/bench_match.py --table > gpurun_out/${tag}_match.txt 2>&1; cat gpurun_out/${tag}_match.txt
timeout 600 python bench.py --steps 20 --warmup 5 --config c4 --no-cpu-baseline > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err
echo "bench c4 rc=$?"; tail -c 300 gpurun_out/${tag}_bench_c4.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_c4.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","steps")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["clocks"])
    print(d["kernels"]["pool_match"]); print(d.get("anchor_match"))
except Exception as e: print("bench parse failed", e)
PY
