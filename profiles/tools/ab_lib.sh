#!/bin/bash
# same-box A/B of two builds of the library: ab_lib.sh <tag> <other .so> [steps]   (alternating runs, two rounds; "new" = in-tree build)
tag=$1; other=$2; steps=${3:-40}
for i in 1 2; do
for v in new old; do
if [ $v = old ]; then export JMT_B200_LIB=$other; else unset JMT_B200_LIB; fi
python bench.py --steps $steps --warmup 5 --gemm-table gpurun_out/gemm_table_${tag}_$v.txt > gpurun_out/bench_${tag}_$v.json 2> gpurun_out/bench_${tag}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${tag}_$v.json"))
print("lib=$v", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["roofline"]["achieved"], d.get("ccc_delta"))
PY
done
done
