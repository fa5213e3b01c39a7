#!/bin/bash
# Builds tuning variants of libpv.so (only pv_gather.cu is recompiled): variants/libpv_<name>.so
# usage: variants.sh name "-DGW_STAGE=256 -DGW_WARPS=2 ..." [name flags ...]
set -e
cd "$(dirname "$0")"
make -s -j8
mkdir -p variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -Xcompiler -fPIC"
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  nvcc $FLAGS $defs -Xptxas -v -c pv_gather.cu -o variants/pv_gather_$name.o 2> variants/$name.ptxas.log
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libpv_$name.so pv_api.o pv_build.o variants/pv_gather_$name.o pv_march.o pv_trace.o pv_shoot.o -lcudart
  echo "$name: $(grep -A2 'gather_kernel' variants/$name.ptxas.log | grep -o 'Used [0-9]* registers' | head -1) $(grep -A1 'properties for _Z13gather_kernel' variants/$name.ptxas.log | tail -1)"
done
