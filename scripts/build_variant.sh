#!/bin/bash
# build_variant.sh NAME "-DFLAG=..." [source.cu ...]: builds imagequilting.jl_b200/libiqb200_NAME.so with the given
# extra nvcc flags applied to the listed sources (default: iq_cutgpu.cu); used with IQB200_LIB=... for kernel experiments.
set -e
NAME=$1; FLAGS=$2; shift 2
SRCS=${@:-iq_cutgpu.cu}
cd "$(dirname "$0")/../imagequilting.jl_b200/csrc"
make -s
OBJS=""
for f in iq_kernels iq_ctx iq_fft iq_cutgpu iq_sim iq_host iq_cut; do
  src=$f.cu; [ -f $src ] || src=$f.cpp
  if echo " $SRCS " | grep -q " $src "; then
    nvcc -gencode arch=compute_100a,code=sm_100a $FLAGS -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-O3,-Wall,-fopenmp -x cu -c $src -o /tmp/${f}_$NAME.o
    OBJS="$OBJS /tmp/${f}_$NAME.o"
  else
    OBJS="$OBJS $f.o"
  fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libiqb200_$NAME.so $OBJS -lgomp
echo built ../libiqb200_$NAME.so
