#!/bin/bash
# same-box A/B of one environment switch: ab_env.sh <tag> <VAR> <value A> <value B> [steps]   (alternating runs, two rounds)
tag=$1; var=$2; a=$3; b=$4; steps=${5:-40}
for i in 1 2; do
for v in $a $b; do
env $var=$v python bench.py --steps $steps --warmup 5 --gemm-table gpurun_out/gemm_table_${tag}_$v.txt > gpurun_out/bench_${tag}_$v.json 2> gpurun_out/bench_${tag}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${tag}_$v.json"))
print("$var=$v", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d.get("ccc_delta"))
PY
done
done
