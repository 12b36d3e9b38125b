#!/bin/bash
tag=${1:-exp4}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_packed_gpu.py -m gpu -q --timeout 300 -x > gpurun_out/${tag}_kernels.log 2>&1; echo "kernels rc=$?"; tail -3 gpurun_out/${tag}_kernels.log
{
for r in 4 3 2; do echo "== MEMVUL_LN_RING_SHORT=$r"; MEMVUL_LN_RING_SHORT=$r timeout 120 python tools/gemm_time.py 2>&1; MEMVUL_LN_RING_SHORT=$r MEMVUL_LN_TRACE=/tmp/ln.bin timeout 120 python tools/ln_trace.py 768 2>&1 | grep "mean cycles"; done
} > gpurun_out/${tag}_micro.txt 2>&1
cat gpurun_out/${tag}_micro.txt
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_configs_gpu.py -m gpu -q --timeout 600 > gpurun_out/${tag}_parity.log 2>&1; echo "parity rc=$?"; tail -4 gpurun_out/${tag}_parity.log
timeout 300 python tools/precision_gpu.py --out gpurun_out/${tag}_precision.json > gpurun_out/${tag}_precision.log 2>&1; echo "precision rc=$?"; grep "^x" gpurun_out/${tag}_precision.log | cut -c1-400
timeout 400 python tools/frontend_bench.py --n 10000 --out gpurun_out/${tag}_frontend.json > gpurun_out/${tag}_frontend.log 2>&1; echo "frontend rc=$?"; tail -2 gpurun_out/${tag}_frontend.log | cut -c1-900
timeout 600 python bench.py --steps 20 --warmup 5 --no-anchor-bench > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","steps")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["clocks"])
    for k,v in d["kernels"].items(): print(k, v.get("avg_us"), v.get("launches_per_step"), v.get("frac_tensor"), v.get("frac_hbm"))
except Exception as e: print("bench parse failed", e)
PY
