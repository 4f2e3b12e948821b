#!/bin/bash
# One GPU visit that produces everything profiles/ is built from (tools/summarize_profiles.py <tag> copies the summaries there):
#   usage: gpurun --timeout 2400 -- 'bash tools/gpu_evidence.sh [quick]'
# quick = tests + bench lines only (no ncu captures)
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/ev_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/ev_pytest_gpu.log
grep -n "passed\|failed\|^FAILED\|^ERROR\|pytest rc" $O/ev_pytest_gpu.log | cut -c1-300 | tail -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/ev_smoke.log 2>&1; echo "smoke rc=$?" >> $O/ev_smoke.log; tail -4 $O/ev_smoke.log | cut -c1-250
timeout 600 python bench.py > $O/ev_bench_default.json 2> $O/ev_bench_default.err; cut -c1-400 $O/ev_bench_default.json; tail -2 $O/ev_bench_default.err | cut -c1-200
timeout 600 python bench.py --impl reference > $O/ev_bench_reference.json 2> $O/ev_bench_reference.err; cut -c1-400 $O/ev_bench_reference.json
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --chunk 32 --batch 16 > $O/ev_bench_chunk32_batch16.json 2> $O/ev_bench_chunk32_batch16.err; cut -c1-300 $O/ev_bench_chunk32_batch16.json
timeout 200 python tools/bench_coarse.py --frames 300 > $O/ev_bench_coarse.json 2> $O/ev_bench_coarse.err; cut -c1-500 $O/ev_bench_coarse.json
timeout 200 python tools/bench_trace.py > $O/ev_bench_trace.json 2> $O/ev_bench_trace.err; cut -c1-300 $O/ev_bench_trace.json
[ "$1" = "quick" ] && exit 0
# launch list of the bench command (per-launch times are cold-cache and serialised: only the kernel's SHARE of a step is comparable)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/ev_ncu_launches.log 2>&1; tail -1 $O/ev_ncu_launches.log | cut -c1-200
# one full capture per hot kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ba_fused_kernel -s 40 -c 1 -f -o $O/prof_ba_fused python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > $O/ev_ncu_ba_fused.log 2>&1; tail -1 $O/ev_ncu_ba_fused.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ba_fused_batch_kernel -s 10 -c 1 -f -o $O/prof_ba_fused_batch python bench.py --steps 10 --warmup 3 --no-cpu-baseline --chunk 32 --batch 16 > $O/ev_ncu_ba_batch.log 2>&1; tail -1 $O/ev_ncu_ba_batch.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ct_track_cluster -s 3 -c 1 -f -o $O/prof_ct_track_cluster python tools/bench_coarse.py --frames 3 --cpu-frames 1 > $O/ev_ncu_ct.log 2>&1; tail -1 $O/ev_ncu_ct.log | cut -c1-200
