#!/bin/bash
# Round-2 profile captures (run on the GPU box through gpurun; summaries: python tools/summarize_profiles.py r2).
# HQS_DEBUG_NO_COOP=1: the tick kernel is launched non-cooperatively so that ncu's kernel replay sees it (148 CTAs of one
# per SM are co-resident either way).
mkdir -p gpurun_out
timeout 600 python bench.py 2> gpurun_out/bench_r2.err | tail -1 > gpurun_out/bench_r2.json
HQS_DEBUG_NO_COOP=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_l.log 2>&1
HQS_DEBUG_NO_COOP=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tick_k -s 6 -c 1 -o gpurun_out/prof_r2 -f python tools/tick_probe.py > gpurun_out/ncu_f.log 2>&1
HQS_DEBUG_NO_COOP=1 timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_prefill.py -k "1000-8-6-2 or 4097-16-12-3 or narrow_amounts_with_remainders or prefill_steal or min_utilization_moves" -x -q > gpurun_out/sanitizer_r2.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_r2.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r2.txt
ls -la gpurun_out | tail -12
