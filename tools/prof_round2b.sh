#!/bin/bash
# r02p: re-capture of the two kernels that changed after the r02n pass (attention: speculative softmax; attn-out GEMM+LN: third residual buffer)
tag=${1:-r02p}
mkdir -p gpurun_out
for k in ln_attn_out:gemm_ln attention:attention_tcgen05; do
  name=${k%%:*}; pat=${k##*:}
  timeout 170 ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 1 -f -o gpurun_out/${tag}_$name python tools/prof_kernels.py $name > gpurun_out/${tag}_ncu_$name.log 2>&1
  echo "$name rc=$?"
done
ls -la gpurun_out | grep ${tag}_ | head
