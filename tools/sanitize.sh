# compute-sanitizer memcheck, bounded by timeouts.  $1 = list of gpu_probe groups (default: the tiny encoder + pool/match)
set -x
for g in ${1:-encoder_tiny poolmatch}; do
  timeout 280 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python tools/gpu_probe.py --child $g > gpurun_out/sanitize_$g.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_$g.log
  grep -c "  ok " gpurun_out/sanitize_$g.log; tail -3 gpurun_out/sanitize_$g.log
done
