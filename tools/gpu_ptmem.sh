#!/bin/bash
# r02q: v3 attention with P handed over through tensor memory (MEMVUL_ATT_PTMEM=1): check + timing, then -- only if the
# check passes -- the attention-bearing GPU tests and a C2 bench with it switched on.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
MEMVUL_ATT_V=3 MEMVUL_ATT_PTMEM=1 timeout 200 python tools/att3_check.py > gpurun_out/ptmem_check.log 2>&1
echo "check rc=$?"; grep -E "FAIL|time|RESULT|rror" gpurun_out/ptmem_check.log | cut -c1-150 | tail -14
grep -q "RESULT.*PASS" gpurun_out/ptmem_check.log || exit 0
export MEMVUL_ATT_PTMEM=1
timeout 300 python -m pytest tests/test_packed_gpu.py tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q --timeout 250 > gpurun_out/ptmem_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/ptmem_tests.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-anchor-bench > gpurun_out/r02q_bench_ptmem.json 2> gpurun_out/r02q_bench_ptmem.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02q_bench_ptmem.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d.get("parity",{}).get("max_logit_err"), d.get("parity",{}).get("ok"), d["clocks"])
    for k in ("attention","attention_cls","gemm_attn_out","gemm_qkv","gemm_ffn_down"): print(" ", k, d["kernels"][k]["avg_us"])
except Exception as e: print("bench parse failed", e)
PY
timeout 200 python -m pytest tests/test_configs_gpu.py -m gpu -q --timeout 180 -k "c2 or c5" > gpurun_out/ptmem_tests_cfg.log 2>&1
echo "config tests rc=$?"; tail -2 gpurun_out/ptmem_tests_cfg.log
