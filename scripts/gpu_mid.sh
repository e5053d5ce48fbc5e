#!/bin/bash
# usage (under gpurun): scripts/gpu_mid.sh [quick] -- mid-kernel tests first, then the whole GPU suite, then C2 A/B and the default line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "mid_kernel" > gpurun_out/pytest_mid.log 2>&1; echo "pytest mid rc=$?"; tail -15 gpurun_out/pytest_mid.log
for v in 2 1 0; do
  if [ $v = 0 ]; then export SB_SOR_MID=0; else export SB_SOR_MID_VARIANT=$v; fi
  SB_MID_TRACE=1 timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c2_v$v.json 2> gpurun_out/bench_c2_v$v.err; echo "c2 variant=$v rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_c2_v$v.json").read().strip().splitlines()[-1])
print(d["config"]["sor_path"], "ms/step", round(d["ms_per_step"],3), "Mcs/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"])
PY
  grep sor_mid_reg gpurun_out/bench_c2_v$v.err | tail -2
  unset SB_SOR_MID SB_SOR_MID_VARIANT
done
[ "$1" = quick ] && exit 0
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
