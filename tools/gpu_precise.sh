#!/bin/bash
# accuracy-mode round: its GPU tests, the throughput comparison, the precision evidence (256 rows x head scales)
tag=${1:-prec}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_precise_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/${tag}_tests.log
timeout 600 python tools/precise_bench.py --out gpurun_out/${tag}_precise_bench.json > gpurun_out/${tag}_precise_bench.log 2>&1; echo "bench rc=$?"; tail -40 gpurun_out/${tag}_precise_bench.log
timeout 900 python tools/precision_gpu.py --precision split_fp16 --rows 256 --out gpurun_out/${tag}_precision_split.json > gpurun_out/${tag}_precision_split.log 2>&1; echo "precision rc=$?"; tail -5 gpurun_out/${tag}_precision_split.log | cut -c1-1500
