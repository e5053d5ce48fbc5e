#!/bin/bash
# usage (under gpurun, one GPU): scripts/gpu_evidence.sh <tag>
# Everything profiles/ quotes for the tree as it is: the default bench line (CPU baseline and
# verification included), one line per BASELINE config at N = 1, the launch list of a tick and
# `ncu --set full` captures of one SOR pass of C5, C3, C4 and of the C2 solve.
TAG=$1
O=gpurun_out
python bench.py > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err
for w in c1 c2 c3 c4; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err
done
timeout 600 python bench.py --mode lex --size 2048 2048 --steps 3 --warmup 1 > $O/${TAG}_bench_lex_2048.json 2> $O/${TAG}_bench_lex_2048.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_tick_8192_T4.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-verify > /dev/null 2>&1
cap() {  # name, kernel regex, skip, count, bench args...
  local n=$1 k=$2 s=$3 c=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -o $O/${TAG}_$n -f \
      python bench.py --steps 1 --warmup 1 --no-cpu --no-verify "$@" > $O/${TAG}_$n.log 2>&1
  # the reports are tens of MB each: keep their digests (gpurun_out/ comes back <= 64 MiB)
  python scripts/ncu_summary.py $O/${TAG}_$n.ncu-rep > $O/${TAG}_${n}_ncu_full.txt 2>&1
  python scripts/ncu_source_digest.py $O/${TAG}_$n.ncu-rep > $O/${TAG}_${n}_stalls.txt 2>&1
  rm -f $O/${TAG}_$n.ncu-rep $O/${TAG}_$n.log
}
cap stream_c5 sor_rb_stream 10 1
cap pass_c3 "sor_rb_(stream_)?kernel" 20 2 --workload c3
cap pass_c4 "sor_rb_(stream_)?kernel" 20 2 --workload c4
cap mid_c2 sor_mid 1 1 --workload c2
cap fg_c5 fg_rhs_fast 2 1
cap adapt_c5 adapt_uv 2 1
for w in default c1 c2 c3 c4 lex_2048; do
  python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_bench_$w.json").read().strip().splitlines()[-1])
    r=d["roofline"] or {}
    print("$w", d["config"]["grid"], "Mcs/s", round(d["value"],2), "ms/step", round(d["ms_per_step"],3), "pass ms", round(r.get("avg_launch_ms",0),4), "e2e", round(d["e2e"]["value"],1), "cpu", (d["cpu_baseline"] or {}).get("value"), "verify", (d["verify"] or {}).get("result"))
except Exception as e:
    print("$w failed", e); print(open("$O/${TAG}_bench_$w.err").read()[-800:])
PY
done
