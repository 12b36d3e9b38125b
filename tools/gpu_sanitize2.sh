#!/bin/bash
# round-2 memcheck pass over the kernels that changed this round (one gpurun call)
mkdir -p gpurun_out
bash tools/sanitize.sh "precise attention gemm_ln" 2>&1 | tail -20
