#!/bin/bash
# fused GEMM+LN (K = 768): paced producer sweep -- time (tools/gemm_time.py, last two lines) and phase trace
for p in 0 600 900 1100 1300 1500; do
  echo "== MEMVUL_LN_PACE=$p"
  MEMVUL_LN_PACE=$p timeout 120 python tools/gemm_time.py 2>&1 | tail -2
  MEMVUL_LN_PACE=$p MEMVUL_LN_TRACE=/tmp/ln.bin timeout 120 python tools/ln_trace.py 768 2>&1 | tail -1
done
