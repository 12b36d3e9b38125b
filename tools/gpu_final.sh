#!/bin/bash
# r02q final pass (one gpurun call): the whole GPU suite + the C2 bench line on the final build, then the evidence for the
# kernel that changed (ncu --set full of the v3 attention, launch list of one step), C5 in data order, memcheck.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
tag=r02q
mkdir -p gpurun_out
bash tools/gpu_round.sh $tag
timeout 120 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05_v3 -s 2 -c 1 -f -o gpurun_out/${tag}_attention python tools/prof_kernels.py attention > gpurun_out/${tag}_ncu_attention.log 2>&1; echo "ncu attention rc=$?"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_raw.csv python bench.py --steps 1 --warmup 1 --preheat-s 0 --no-cpu-baseline --no-anchor-bench > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 120 python bench.py --config c5 --steps 20 --warmup 5 --no-cpu-baseline --no-anchor-bench > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err; echo "bench c5 rc=$?"; tail -c 300 gpurun_out/${tag}_bench_c5.json | head -c 300; echo
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python tools/gpu_probe.py --child attention > gpurun_out/${tag}_sanitize_attention.log 2>&1; echo "memcheck rc=$?"; grep -c "  ok " gpurun_out/${tag}_sanitize_attention.log; tail -3 gpurun_out/${tag}_sanitize_attention.log
ls -la gpurun_out | grep ${tag}_ | wc -l
