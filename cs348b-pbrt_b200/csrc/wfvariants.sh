#!/bin/bash
# Build variants of the wavefront shooter (pv_wavefront.cu) for A/B runs on the GPU box: usage ./wfvariants.sh name "-DX=.." [name defs ...]
# -> variants/libpv_<name>.so, selected at run time with PV_LIBPV=<path>.  variants/ is git-ignored but travels with the snapshot.
cd "$(dirname "$0")"
mkdir -p variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -Xcompiler -fPIC -Xcompiler -Wno-unused-function"
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  nvcc $FLAGS $defs -Xptxas -v -c pv_wavefront.cu -o variants/pv_wavefront_$name.o 2> variants/$name.wf.ptxas.log
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libpv_$name.so pv_api.o pv_build.o pv_gather.o pv_cellgather.o pv_comm.o pv_march.o pv_trace.o pv_shoot.o variants/pv_wavefront_$name.o pv_volint.o -lcudart -ldl
  echo "$name: $(grep -A2 'wf_march_kernelILi0' variants/$name.wf.ptxas.log | grep -o 'Used [0-9]* registers' | head -1) $(grep -A1 'properties for _Z15wf_march_kernelILi0' variants/$name.wf.ptxas.log | tail -1 | xargs)"
done
