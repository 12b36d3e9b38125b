#!/bin/bash
# r02q: attention v4 (two query tiles per CTA, token ping-pong) against v3 -- correctness + timing, then in-step
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for cfg in "3 1" "4 1" "4 0"; do
  set -- $cfg
  MEMVUL_ATT_V=$1 MEMVUL_ATT4_TOKEN=$2 timeout 240 python tools/att3_check.py > gpurun_out/att4_v$1_t$2.log 2>&1
  echo "exit $? (v=$1 token=$2)"; grep -E "FAIL|time|RESULT|rror" gpurun_out/att4_v$1_t$2.log | cut -c1-160 | tail -12
done
if grep -q "RESULT.*PASS" gpurun_out/att4_v4_t1.log; then
  MEMVUL_ATT_V=4 timeout 300 python -m pytest tests/test_packed_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout 250 > gpurun_out/att4_tests.log 2>&1
  echo "tests(v4) rc=$?"; tail -3 gpurun_out/att4_tests.log
  MEMVUL_ATT_V=4 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-anchor-bench > gpurun_out/r02q_bench_attv4.json 2> gpurun_out/r02q_bench_attv4.err
  echo "bench v4 rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02q_bench_attv4.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d.get("parity",{}).get("max_logit_err"), d["clocks"])
    for k in ("attention","attention_cls","gemm_attn_out","gemm_qkv"): print(" ", k, d["kernels"][k]["kernel"], d["kernels"][k]["avg_us"])
except Exception as e: print("bench parse failed", e)
PY
fi
