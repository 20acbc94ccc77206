#!/bin/bash
# usage: ncu_gemm.sh tag M N K cluster
tag=$1; M=$2; N=$3; K=$4; cl=$5
JMT_GEMM_CLUSTER=$cl ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip 2 -c 1 -f -o gpurun_out/ncu_$tag python profiles/tools/gemm_one.py $M $N $K 4 > gpurun_out/ncu_$tag.log 2>&1
ncu -i gpurun_out/ncu_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_$tag.raw.csv 2>/dev/null
