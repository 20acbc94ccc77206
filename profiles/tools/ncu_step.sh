#!/bin/bash
# launch list of one eager step + full capture of the dominant kernels inside the real step
tag=$1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1900 -c 380 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --no-graph > gpurun_out/ncu_launch_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_chain --launch-skip 40 -c 1 -f -o gpurun_out/ncu_attn_$tag python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --no-graph > gpurun_out/ncu_attn_$tag.log 2>&1
ncu -i gpurun_out/ncu_attn_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_attn_$tag.raw.csv 2>/dev/null
