#!/bin/bash
# r02q: attention v3 (three CTAs per SM) against v1 -- correctness + timing, one process per variant
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for cfg in "1 0" "3 0" "3 2"; do
  set -- $cfg
  MEMVUL_ATT_V=$1 MEMVUL_ATT_POLY=$2 timeout 240 python tools/att3_check.py > gpurun_out/att3_v$1_p$2.log 2>&1
  echo "exit $? (v=$1 poly=$2)"; tail -n 5 gpurun_out/att3_v$1_p$2.log | cut -c1-160
done
