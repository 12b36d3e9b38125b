#!/bin/bash
tag=${1:-exp2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_packed_gpu.py -m gpu -q --timeout 300 -x > gpurun_out/${tag}_kernels.log 2>&1; echo "kernels rc=$?"; tail -5 gpurun_out/${tag}_kernels.log
{
for v in 1 2; do echo "== MEMVUL_ATT_V=$v"; MEMVUL_ATT_V=$v timeout 120 python tools/att_time.py; done
MEMVUL_ATT_TRACE=/tmp/att2.bin timeout 120 python tools/att_trace2.py
for mode in 0 2 6; do
  echo "== MEMVUL_LN_MODE=$mode"
  MEMVUL_LN_MODE=$mode MEMVUL_LN_TRACE=/tmp/ln.bin timeout 120 python tools/ln_trace.py 768 2>&1 | grep "mean cycles"
  MEMVUL_LN_MODE=$mode timeout 120 python tools/gemm_time.py 2>&1 | grep "gemm_ln"
done
} > gpurun_out/${tag}_micro.txt 2>&1
cat gpurun_out/${tag}_micro.txt
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_configs_gpu.py -m gpu -q --timeout 600 -x > gpurun_out/${tag}_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/${tag}_parity.log
for mode in 0 6; do
MEMVUL_LN_MODE=$mode timeout 600 python bench.py --steps 20 --warmup 5 --no-anchor-bench --no-cpu-baseline > gpurun_out/${tag}_bench_m$mode.json 2> gpurun_out/${tag}_bench_m$mode.err
echo "bench mode $mode rc=$?"; tail -c 300 gpurun_out/${tag}_bench_m$mode.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_m$mode.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","steps")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["clocks"])
    for k,v in d["kernels"].items(): print(k, v.get("avg_us"), v.get("launches_per_step"), v.get("frac_tensor"), v.get("frac_hbm"))
except Exception as e: print("bench parse failed", e)
PY
done
