#!/bin/bash
# diagnostic build of the library with per-phase cycle counters in the GEMM epilogue (epilogue warp 0 of CTA 0)
set -e
cd "$(dirname "$0")/../.."
PKG=joint-multimodal-transformer-6th-abaw_b200
mkdir -p profiles/tools/ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -cudart static -DJMT_EPI_PROF -c $PKG/csrc/gemm_tc.cu -o profiles/tools/ab/gemm_tc_prof.o
objs=$(ls $PKG/build/*.o | grep -v gemm_tc.cu.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o profiles/tools/ab/lib_prof.so $objs profiles/tools/ab/gemm_tc_prof.o
echo profiles/tools/ab/lib_prof.so
