#!/bin/bash
# tools/build_variant.sh NAME "-DMJB_CTAS_INERTIA=3 ..."  ->  mujoco_inversedynamicstest_b200/lib/variants/libmjb_NAME.so
# (tuning experiments: select with MJB_LIB=... on the GPU box; the shipped library is lib/libmjb.so)
set -e
name=$1; defs=$2
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/mujoco_inversedynamicstest_b200/csrc
out=$root/mujoco_inversedynamicstest_b200/lib/variants
obj=$root/build/variants/$name
mkdir -p $out $obj
inc=${MUJOCO_INCLUDE:-/root/reference/include}
flags="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -I$inc -I$root/include -I$src $defs"
nvcc $flags -Xptxas -v -c $src/mjb_kernels.cu -o $obj/mjb_kernels.o 2> $obj/ptxas.log &
for f in mjb_api.cu mjb_upload.cc mjb_modelio.cc; do nvcc $flags -c $src/$f -o $obj/${f%.*}.o & done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libmjb_$name.so $obj/*.o -cudart static
grep -A1 "Compiling entry function.*kernelILb1" $obj/ptxas.log | grep -v "^--" | paste - - | sed 's/.*_Z[0-9N]*3mjb[0-9]*\([a-z_]*\)ILb1.*Used \([0-9]*\) registers.*/\1 \2 regs/' | sort -u
