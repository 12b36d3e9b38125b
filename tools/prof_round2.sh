#!/bin/bash
# r02 evidence pass (one gpurun call): `ncu --set full` of every hot kernel at the C2 shape, of the anchor match in its
# three regimes, and the launch list of one bench step.  Reports land in gpurun_out/<tag>_*.ncu-rep; summaries are made
# with tools/ncu_summary.py in the build container and committed under profiles/.
tag=${1:-r02n}
mkdir -p gpurun_out
for k in gemm_qkv:gemm_f16_tcgen05_2cta gemm_ffn_up:gemm_f16_tcgen05_2cta ln_attn_out:gemm_ln ln_ffn_down:gemm_ln attention:attention_tcgen05 pool_match:pool_match; do
  name=${k%%:*}; pat=${k##*:}
  timeout 170 ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 1 -f -o gpurun_out/${tag}_$name python tools/prof_kernels.py $name > gpurun_out/${tag}_ncu_$name.log 2>&1
  echo "$name rc=$?"
done
PM_B=1 PM_G=262144 timeout 170 ncu --set full --clock-control none -k regex:pool_match -s 2 -c 1 -f -o gpurun_out/${tag}_pool_match_streaming python tools/prof_kernels.py pool_match > gpurun_out/${tag}_ncu_pm_stream.log 2>&1; echo "pm streaming rc=$?"
PM_B=256 PM_G=16384 timeout 170 ncu --set full --clock-control none -k regex:pool_match -s 2 -c 1 -f -o gpurun_out/${tag}_pool_match_c4 python tools/prof_kernels.py pool_match > gpurun_out/${tag}_ncu_pm_c4.log 2>&1; echo "pm c4 rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_raw.csv python bench.py --steps 1 --warmup 1 --preheat-s 0 --no-cpu-baseline --no-anchor-bench > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out | grep ${tag} | head -20
