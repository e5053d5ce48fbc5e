#!/bin/bash
# usage (under gpurun): scripts/gpu_round.sh  -- GPU tests, smoke, bench lines, ncu launch list + full captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cat gpurun_out/bench_default.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cat gpurun_out/bench_reference.json
timeout 300 python bench.py --steps 10 --warmup 3 --tblock 3 --no-cpu > gpurun_out/bench_T3.json 2> gpurun_out/bench_T3.err; cat gpurun_out/bench_T3.json
timeout 300 python bench.py --steps 3 --warmup 1 --mode lex --no-cpu --size 2048 2048 > gpurun_out/bench_lex_2048.json 2> gpurun_out/bench_lex_2048.err; cat gpurun_out/bench_lex_2048.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tick_8192_T4.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sor_rb_stream -s 6 -c 1 -o gpurun_out/stream_T4_full -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fg_rhs_kernel -s 1 -c 1 -o gpurun_out/fg_rhs_full -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full3.log 2>&1
scripts/all_workloads.sh
ls gpurun_out | head -50
