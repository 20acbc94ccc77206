#!/bin/bash
# (attn_chain_kernel launches per eager step: 9 forward (mode 0) then 9 backward (mode 1); 5 untimed steps = 90 launches come first)
# Round-2 evidence: launch list of one eager C2 step + `ncu --set full` of the dominant GEMM shape and of the fused attention kernels
# inside the real step.   usage: bash profiles/tools/ncu_round2.sh <tag>      (run under gpurun; never a bench value)
tag=$1
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --no-graph --no-parity"
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 900 -c 800 --csv --log-file gpurun_out/launches_${tag}_raw.csv $B > gpurun_out/ncu_launch_$tag.log 2>&1
python profiles/tools/launch_step.py gpurun_out/launches_${tag}_raw.csv gpurun_out/launches_${tag}_eager_step.csv > gpurun_out/launches_${tag}_summary.txt
ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip 2 -c 1 -f -o gpurun_out/ncu_gemm_$tag python profiles/tools/gemm_one.py 76800 512 512 4 > gpurun_out/ncu_gemm_$tag.log 2>&1
ncu -i gpurun_out/ncu_gemm_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_gemm_$tag.raw.csv 2>/dev/null
python profiles/tools/ncu_extract.py gpurun_out/ncu_gemm_$tag.raw.csv > gpurun_out/ncu_gemm_tc_${tag}_linear_76800x512x512.csv
ncu --set full --clock-control none --import-source on -k regex:attn_chain_kernel --launch-skip 90 -c 1 -f -o gpurun_out/ncu_attn_fwd_$tag $B > gpurun_out/ncu_attn_fwd_$tag.log 2>&1
ncu -i gpurun_out/ncu_attn_fwd_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_attn_fwd_$tag.raw.csv 2>/dev/null
python profiles/tools/ncu_extract.py gpurun_out/ncu_attn_fwd_$tag.raw.csv > gpurun_out/ncu_attn_chain_${tag}_fwd_300x300x512_nb256.csv
ncu --set full --clock-control none --import-source on -k regex:attn_chain_kernel --launch-skip 99 -c 1 -f -o gpurun_out/ncu_attn_bwd_$tag $B > gpurun_out/ncu_attn_bwd_$tag.log 2>&1
ncu -i gpurun_out/ncu_attn_bwd_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_attn_bwd_$tag.raw.csv 2>/dev/null
python profiles/tools/ncu_extract.py gpurun_out/ncu_attn_bwd_$tag.raw.csv > gpurun_out/ncu_attn_chain_${tag}_bwd_ds_300x300x512_nb256.csv
rm -f gpurun_out/ncu_*_$tag.ncu-rep
