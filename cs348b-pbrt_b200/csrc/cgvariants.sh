#!/bin/bash
# Builds tuning variants of libpv.so that differ in pv_cellgather.cu / pv_gather.cu only: variants/libpv_<name>.so.
# usage: cgvariants.sh name "<cellgather defs>" [name defs ...]
set -e
cd "$(dirname "$0")"
make -s -j8
mkdir -p variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -Xcompiler -fPIC"
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  nvcc $FLAGS $defs -Xptxas -v -c pv_cellgather.cu -o variants/pv_cellgather_$name.o 2> variants/$name.cg.ptxas.log
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libpv_$name.so pv_api.o pv_build.o pv_gather.o variants/pv_cellgather_$name.o pv_comm.o pv_march.o pv_trace.o pv_shoot.o pv_wavefront.o pv_volint.o -lcudart -ldl
  echo "$name: $(grep -A2 'cellgather_kernel' variants/$name.cg.ptxas.log | grep -o 'Used [0-9]* registers' | head -1) $(grep -A1 'properties for _Z17cellgather' variants/$name.cg.ptxas.log | tail -1 | xargs)"
done
