mkdir -p gpurun_out
timeout 500 python bench.py 2>&1 | grep -v transformNew | tail -1 > gpurun_out/bench_r1.json
HQS_DEBUG_NO_COOP=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-drain > gpurun_out/ncu_l.log 2>&1
HQS_DEBUG_NO_COOP=1 timeout 500 ncu --set full --clock-control none --import-source on -k "regex:count_k|solve_k|emit_k" -s 27 -c 3 -o gpurun_out/prof_r1 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-drain > gpurun_out/ncu_f.log 2>&1
HQS_DEBUG_NO_COOP=1 HQS_DEBUG_SMALL_GRID=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:solve_k -s 4 -c 1 -o gpurun_out/prof_r1_solve -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-drain > gpurun_out/ncu_s.log 2>&1
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | grep -v transformNew | tail -1 > gpurun_out/bench_ref_r1.json
ls gpurun_out | head -30
