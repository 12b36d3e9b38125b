#!/bin/bash
# fused GEMM+LN: third residual buffer (MEMVUL_LN_XBUF) on / off -- parity, time (tools/gemm_time.py), phase trace
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "layernorm_fused" 2>&1 | tail -3
for x in 0 1; do
  echo "== MEMVUL_LN_XBUF=$x"
  MEMVUL_LN_XBUF=$x timeout 120 python tools/gemm_time.py 2>&1 | tail -2
  MEMVUL_LN_XBUF=$x MEMVUL_LN_TRACE=/tmp/ln.bin timeout 120 python tools/ln_trace.py 768 2>&1 | tail -1
done
