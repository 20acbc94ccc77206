#!/bin/bash
# ncu evidence for the memory-bound kernels (HBM GB/s) and for the attention backward (dS) kernel, inside a real eager step
tag=$1
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --no-graph"
timeout 420 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none \
  -k regex:'add_layernorm|colsum_kernel|act_bwd|add_act|l2norm|ccc_sums|ccc_bwd|time_max|regressor_tail|weight_norm|transpose_kernel|copy_rows3d|copy2d|rowdot|dropout_mask|pad_right' \
  --launch-skip 330 -c 130 -f -o gpurun_out/ncu_membound_$tag $B > gpurun_out/ncu_membound_$tag.log 2>&1
ncu -i gpurun_out/ncu_membound_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_membound_$tag.raw.csv 2>/dev/null
python profiles/tools/ncu_membound_summary.py gpurun_out/ncu_membound_$tag.raw.csv > gpurun_out/ncu_membound_$tag.summary.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_chain --launch-skip 13 -c 1 -f -o gpurun_out/ncu_attn_ds_$tag $B > gpurun_out/ncu_attn_ds_$tag.log 2>&1
ncu -i gpurun_out/ncu_attn_ds_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_attn_ds_$tag.raw.csv 2>/dev/null
python profiles/tools/ncu_extract.py gpurun_out/ncu_attn_ds_$tag.raw.csv > gpurun_out/ncu_attn_ds_$tag.summary.csv
