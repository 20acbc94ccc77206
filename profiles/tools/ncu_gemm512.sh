tag=r3p
ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip 2 -c 1 -f -o gpurun_out/ncu_gemm_$tag python profiles/tools/gemm_one.py 76800 512 512 4 > gpurun_out/ncu_gemm_$tag.log 2>&1
ncu -i gpurun_out/ncu_gemm_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_gemm_$tag.raw.csv 2>/dev/null
python profiles/tools/ncu_extract.py gpurun_out/ncu_gemm_$tag.raw.csv > gpurun_out/ncu_gemm_tc_${tag}_linear_76800x512x512.csv
rm -f gpurun_out/ncu_gemm_$tag.ncu-rep gpurun_out/ncu_gemm_$tag.raw.csv
cat gpurun_out/ncu_gemm_tc_${tag}_linear_76800x512x512.csv
