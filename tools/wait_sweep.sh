for v in "5 5" "21 21" "42 42" "5 5" "21 21" "0 0"; do set -- $v
  MEMVUL_GEMM_WAIT=$1 MEMVUL_ATT_WAIT=$2 timeout 200 python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('WAIT=$1/$2', round(d['value'],1), 'issues/s', round(d['ms_per_step'],3), 'ms', d['clocks']['sm_mhz'], 'MHz', 'e2e', round(d['e2e']['value'],1))"
done
