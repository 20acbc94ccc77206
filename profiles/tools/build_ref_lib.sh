#!/bin/bash
# build the library of another git revision next to the working tree (same-box A/B through JMT_B200_LIB):
#   bash profiles/tools/build_ref_lib.sh <rev> <out.so>
set -e
rev=$1; out=$2
tmp=$(mktemp -d)
git archive $rev joint-multimodal-transformer-6th-abaw_b200/csrc include | tar -x -C $tmp
cd $tmp/joint-multimodal-transformer-6th-abaw_b200/csrc
objs=""
for f in *.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -cudart static -c $f -o $f.o 2>/dev/null &
  objs="$objs $f.o"
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o $out $objs
cd /; rm -rf $tmp
echo $out
