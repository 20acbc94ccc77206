#!/bin/bash
# Round-2 (second half) evidence: `ncu --set full` of the dominant GEMM shape, of the fused attention kernels and of the dQ/dK/dV kernel
# inside the real step (launch list: ncu_launches.sh).   usage: bash profiles/tools/ncu_round3.sh <tag>   (under gpurun; never a bench value)
tag=$1
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --no-graph --no-parity"
ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip 2 -c 1 -f -o gpurun_out/ncu_gemm_$tag python profiles/tools/gemm_one.py 76800 512 512 4 > gpurun_out/ncu_gemm_$tag.log 2>&1
ncu -i gpurun_out/ncu_gemm_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_gemm_$tag.raw.csv 2>/dev/null
python profiles/tools/ncu_extract.py gpurun_out/ncu_gemm_$tag.raw.csv > gpurun_out/ncu_gemm_tc_${tag}_linear_76800x512x512.csv
ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip 2 -c 1 -f -o gpurun_out/ncu_gemm3072_$tag python profiles/tools/gemm_one.py 76800 1024 3072 4 > gpurun_out/ncu_gemm3072_$tag.log 2>&1
ncu -i gpurun_out/ncu_gemm3072_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_gemm3072_$tag.raw.csv 2>/dev/null
python profiles/tools/ncu_extract.py gpurun_out/ncu_gemm3072_$tag.raw.csv > gpurun_out/ncu_gemm_tc_${tag}_linear_76800x1024x3072.csv
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_dqkv --launch-skip 30 -c 1 -f -o gpurun_out/ncu_dqkv_$tag $B > gpurun_out/ncu_dqkv_$tag.log 2>&1
ncu -i gpurun_out/ncu_dqkv_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_dqkv_$tag.raw.csv 2>/dev/null
python profiles/tools/ncu_extract.py gpurun_out/ncu_dqkv_$tag.raw.csv > gpurun_out/ncu_attn_bwd_dqkv_${tag}_300x300x512_nb256.csv
ncu --set full --clock-control none --import-source on -k regex:attn_chain_kernel --launch-skip 63 -c 1 -f -o gpurun_out/ncu_attn_fwd_$tag $B > gpurun_out/ncu_attn_fwd_$tag.log 2>&1
ncu -i gpurun_out/ncu_attn_fwd_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_attn_fwd_$tag.raw.csv 2>/dev/null
python profiles/tools/ncu_extract.py gpurun_out/ncu_attn_fwd_$tag.raw.csv > gpurun_out/ncu_attn_chain_${tag}_300x300x512_nb256.csv
rm -f gpurun_out/ncu_*_$tag.ncu-rep gpurun_out/ncu_*_$tag.raw.csv
head -40 gpurun_out/ncu_attn_bwd_dqkv_${tag}_300x300x512_nb256.csv
grep -E "Kernel Name|time_duration|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|dram__bytes" gpurun_out/ncu_gemm_tc_${tag}_linear_76800x512x512.csv gpurun_out/ncu_gemm_tc_${tag}_linear_76800x1024x3072.csv gpurun_out/ncu_attn_chain_${tag}_300x300x512_nb256.csv
