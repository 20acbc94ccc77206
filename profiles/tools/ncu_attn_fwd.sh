#!/bin/bash
# `ncu --set full` of the fused attention forward kernel inside the real step (never a bench value):  bash profiles/tools/ncu_attn_fwd.sh <tag>
tag=$1
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --no-graph --no-parity"
ncu --set full --clock-control none --import-source on -k regex:attn_chain_kernel --launch-skip 56 -c 1 -f -o gpurun_out/ncu_attn_fwd_$tag $B > gpurun_out/ncu_attn_fwd_$tag.log 2>&1
ncu -i gpurun_out/ncu_attn_fwd_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_attn_fwd_$tag.raw.csv 2>/dev/null
python profiles/tools/ncu_extract.py gpurun_out/ncu_attn_fwd_$tag.raw.csv > gpurun_out/ncu_attn_chain_${tag}_fwd_300x300x512_nb256.csv
rm -f gpurun_out/ncu_attn_fwd_$tag.ncu-rep gpurun_out/ncu_attn_fwd_$tag.raw.csv
cat gpurun_out/ncu_attn_chain_${tag}_fwd_300x300x512_nb256.csv
