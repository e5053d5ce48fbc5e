#!/bin/bash
# usage (under gpurun): scripts/gpu_round.sh  -- GPU tests, smoke, bench lines, ncu launch list + full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
for T in 3 4; do
  timeout 300 python bench.py --steps 10 --warmup 3 --tblock $T > gpurun_out/bench_T$T.json 2> gpurun_out/bench_T$T.err; echo "bench T=$T rc=$?"
  cat gpurun_out/bench_T$T.json
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tick_8192_T3.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sor_rb_stream -s 10 -c 1 -o gpurun_out/stream_T3_full -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
