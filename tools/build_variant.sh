#!/bin/bash
# tools/build_variant.sh NAME "-DMJB_INDEX_GLOBAL=1 ..."  ->  mujoco_inversedynamicstest_b200/lib/variants/libmjb_NAME.so
# Tuning experiments: mjb_kernels.cu recompiled with the given defines and linked with the objects of the
# shipped build (run build() first); select with MJB_LIB=... on the GPU box. The shipped library is lib/libmjb.so.
set -e
name=$1; defs=$2
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/mujoco_inversedynamicstest_b200/csrc
out=$root/mujoco_inversedynamicstest_b200/lib/variants
base=$root/build/mjb
obj=$root/build/variants/$name
mkdir -p $out $obj
inc=${MUJOCO_INCLUDE:-/root/reference/include}
flags="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden -I$inc -I$root/include -I$src -I$base $defs"
nvcc $flags -Xptxas -v -c $src/mjb_kernels.cu -o $obj/mjb_kernels.o 2> $obj/ptxas.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libmjb_$name.so $obj/mjb_kernels.o $base/mjb_api.o $base/mjb_jit.o $base/mjb_upload.o $base/mjb_modelio.o -cudart static -ldl
grep -A1 "Compiling entry function.*\(rows\|index\|narrow\)_kernelILb" $obj/ptxas.log | grep -o "mjb[0-9]*[a-z_]*kernelILb[01]ELb[01]\|mjb[0-9]*[a-z_]*kernelILb[01]\|Used [0-9]* registers" | paste - - | sort -u
