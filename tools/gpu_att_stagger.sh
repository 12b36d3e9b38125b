#!/bin/bash
# attention variants: timing (tools/att_time.py), phase trace (tools/att_trace.py), parity of the attention / packed tests
run() { echo "== $*"; env "$@" MEMVUL_ATT_TRACE=/tmp/att.bin timeout 120 python tools/att_trace.py 2>&1 | head -9; env "$@" timeout 120 python tools/att_time.py 2>&1 | head -4; }
run MEMVUL_ATT_SPEC=0
run MEMVUL_ATT_SPEC=1
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_packed_gpu.py -m gpu -q -k "attention or packed" 2>&1 | tail -3
