#!/bin/bash
tag=$1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1750 -c 350 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --no-graph > gpurun_out/ncu_launch_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_chain_kernel.*1 --launch-skip 9 -c 1 -f -o gpurun_out/ncu_attn_ds_$tag python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --no-graph > gpurun_out/ncu_attn_ds_$tag.log 2>&1
ncu -i gpurun_out/ncu_attn_ds_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_attn_ds_$tag.raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel --launch-skip 400 -c 3 -f -o gpurun_out/ncu_gemm_step_$tag python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --no-graph > gpurun_out/ncu_gemm_step_$tag.log 2>&1
ncu -i gpurun_out/ncu_gemm_step_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_gemm_step_$tag.raw.csv 2>/dev/null
