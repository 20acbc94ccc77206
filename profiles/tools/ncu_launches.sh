#!/bin/bash
# launch list of one eager C2 step (never a bench value):  bash profiles/tools/ncu_launches.sh <tag>
tag=$1
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --no-graph --no-parity"
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 800 -c 800 --csv --log-file gpurun_out/launches_${tag}_raw.csv $B > gpurun_out/ncu_launch_$tag.log 2>&1
python profiles/tools/launch_step.py gpurun_out/launches_${tag}_raw.csv gpurun_out/launches_${tag}_eager_step.csv > gpurun_out/launches_${tag}_summary.txt
cat gpurun_out/launches_${tag}_summary.txt
