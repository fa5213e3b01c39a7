#!/bin/bash
# Builds tuning variants of libpv.so: variants/libpv_<name>.so with extra -D flags for pv_gather.cu / pv_march.cu.
# usage: [SDEFS="<shoot defs>"] variants.sh name "<gather defs>" "<march defs>" [name gdefs mdefs ...]
set -e
cd "$(dirname "$0")"
make -s -j8
mkdir -p variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -Xcompiler -fPIC"
while [ $# -ge 3 ]; do
  name=$1; gdefs=$2; mdefs=$3; shift 3
  nvcc $FLAGS $gdefs -Xptxas -v -c pv_gather.cu -o variants/pv_gather_$name.o 2> variants/$name.g.ptxas.log &
  nvcc $FLAGS $mdefs -Xptxas -v -c pv_march.cu -o variants/pv_march_$name.o 2> variants/$name.m.ptxas.log &
  nvcc $FLAGS $SDEFS -Xptxas -v -c pv_shoot.cu -o variants/pv_shoot_$name.o 2> variants/$name.s.ptxas.log &
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libpv_$name.so pv_api.o pv_build.o variants/pv_gather_$name.o variants/pv_march_$name.o pv_trace.o variants/pv_shoot_$name.o pv_volint.o -lcudart
  echo "$name: shoot $(grep -A3 'shoot_kernel' variants/$name.s.ptxas.log | grep -o 'Used [0-9]* registers' | head -1), $(grep -A1 'properties for _Z12shoot_kernel' variants/$name.s.ptxas.log | tail -1 | xargs)"
  echo "$name: gather $(grep -A2 'gather_kernel' variants/$name.g.ptxas.log | grep -o 'Used [0-9]* registers' | head -1), $(grep -A1 'properties for _Z13gather_kernel' variants/$name.g.ptxas.log | tail -1 | xargs) | march $(grep -A2 'march_steps' variants/$name.m.ptxas.log | grep -o 'Used [0-9]* registers' | head -1), $(grep -A1 'properties for _Z18march_steps' variants/$name.m.ptxas.log | tail -1 | xargs)"
done
