#!/bin/bash
# r02q: in-step A/B of the attention kernels (v1 / v3), the 8 x 7 register-tile match against the 64 x 64 one, and the
# tests that cover both changes.  One gpurun call.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for v in 1 3; do
  MEMVUL_ATT_V=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-anchor-bench > gpurun_out/r02q_bench_attv$v.json 2> gpurun_out/r02q_bench_attv$v.err
  echo "bench v$v rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02q_bench_attv$v.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d.get("parity",{}).get("max_logit_err"), d["clocks"])
    for k in ("attention","gemm_attn_out","gemm_qkv","gemm_ffn_up","gemm_ffn_down"): print(" ", k, d["kernels"][k]["kernel"], d["kernels"][k]["avg_us"])
except Exception as e: print("bench parse failed", e)
PY
done
for t in "tests/test_configs_gpu.py -k c4" "tests/test_kernels_gpu.py" "tests/test_packed_gpu.py"; do
  n=$(echo $t | cut -d/ -f2 | cut -d. -f1)
  timeout 600 python -m pytest $t -m gpu -q --timeout 500 > gpurun_out/r02q_$n.log 2>&1
  echo "$n rc=$?"; tail -3 gpurun_out/r02q_$n.log
done
MEMVUL_MATCH_TILED=1 timeout 200 python tools/bench_match.py --c4 --table > gpurun_out/r02q_match_tile64.txt 2>&1; cat gpurun_out/r02q_match_tile64.txt
timeout 200 python tools/bench_match.py --c4 --table > gpurun_out/r02q_match_tile8.txt 2>&1; cat gpurun_out/r02q_match_tile8.txt
