#!/bin/bash
# multi-GPU validation (gpurun --gpus N): NCCL gather test, second-device test, scaling bench lines
tag=${1:-multi}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest "tests/test_configs_gpu.py::test_nccl_gather_match_two_ranks" "tests/test_packed_gpu.py::test_model_on_a_second_device_while_device0_is_current" -m gpu -q --timeout 300 > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${tag}_tests.log
for cfg in c2 c5; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --config $cfg > gpurun_out/${tag}_bench_${cfg}_n$N.json 2> gpurun_out/${tag}_bench_${cfg}_n$N.err
echo "bench $cfg N=$N rc=$?"; tail -c 500 gpurun_out/${tag}_bench_${cfg}_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_${cfg}_n$N.json").read().strip().splitlines()[-1])
    print("$cfg", {k:d[k] for k in ("value","ms_per_step","steps","n_gpus")}, "e2e", d["e2e"]["value"], d["config"]["global_batch"])
except Exception as e: print("bench parse failed", e)
PY
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-anchor-bench --no-cpu-baseline --config c2 > gpurun_out/${tag}_bench_c2_n1.json 2>/dev/null
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_c2_n1.json").read().strip().splitlines()[-1])
    print("c2 n1", {k:d[k] for k in ("value","ms_per_step","steps","n_gpus")}, "e2e", d["e2e"]["value"])
except Exception as e: print("bench parse failed", e)
PY
