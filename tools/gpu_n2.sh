#!/bin/bash
# 2-GPU round: the NCCL tests (skipped on one GPU) and the C2 bench line at N=2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_configs_gpu.py tests/test_packed_gpu.py -m gpu -q --timeout 500 -k "nccl or second_device" > gpurun_out/n2_nccl_tests.log 2>&1; echo "nccl tests rc=$?"; tail -4 gpurun_out/n2_nccl_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_bench_c2.json 2> gpurun_out/n2_bench_c2.err; echo "bench rc=$?"; tail -c 300 gpurun_out/n2_bench_c2.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/n2_bench_c2.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","n_gpus","ms_per_step","steps")}, "e2e", d["e2e"]["value"], d["clocks"])
except Exception as e: print("parse failed", e)
PY
