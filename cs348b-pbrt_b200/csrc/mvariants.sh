#!/bin/bash
# Build variants of the march kernels (pv_march.cu): usage ./mvariants.sh name "-DX=.." [name defs ...] -> variants/libpv_<name>.so
cd "$(dirname "$0")"
mkdir -p variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -Xcompiler -fPIC -Xcompiler -Wno-unused-function"
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  nvcc $FLAGS $defs -Xptxas -v -c pv_march.cu -o variants/pv_march_$name.o 2> variants/$name.m.ptxas.log
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libpv_$name.so pv_api.o pv_build.o pv_gather.o pv_cellgather.o pv_comm.o variants/pv_march_$name.o pv_trace.o pv_shoot.o pv_wavefront.o pv_volint.o -lcudart -ldl
  echo "$name: $(grep -A2 'march_steps_kernelILb0' variants/$name.m.ptxas.log | grep -o 'Used [0-9]* registers' | head -1)"
done
