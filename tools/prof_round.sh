set -x
for k in gemm_qkv:gemm_f16_tcgen05_2cta gemm_ffn_up:gemm_f16_tcgen05_2cta ln_attn_out:gemm_ln ln_ffn_down:gemm_ln attention:attention_tcgen05; do
  name=${k%%:*}; pat=${k##*:}
  timeout 170 ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 1 -f -o gpurun_out/r01n_$name python tools/prof_kernels.py $name > gpurun_out/ncu_$name.log 2>&1
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 194 -c 162 --csv --log-file gpurun_out/launches_r1n.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out | tail -8
