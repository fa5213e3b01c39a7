#!/bin/bash
# What the first GPU call of the next round should run (nothing here has been run on hardware yet): the two pieces written
# without a GPU at the end of round 1.   gpurun --timeout 900 -- 'bash tools/next_gpu_call.sh'      (part 2 needs --gpus 2)
cd /root/repo
mkdir -p gpurun_out
# 1. GW_RANGES_PRE (cell ranges from a thread-per-step kernel): the variant library was built here by
#    `cs348b-pbrt_b200/csrc/variants.sh rpre "-DGW_RANGES_PRE=1" ""` (variants/ is git-ignored but travels with the snapshot).
V=/root/repo/cs348b-pbrt_b200/csrc/variants/libpv_rpre.so
[ -f $V ] || bash cs348b-pbrt_b200/csrc/variants.sh rpre "-DGW_RANGES_PRE=1" "" > gpurun_out/rpre_build.log 2>&1   # better done before the call
if [ -f $V ]; then
  PV_LIBPV=$V python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/rpre_parity.log 2>&1
  tail -3 gpurun_out/rpre_parity.log
  export PV_BENCH_CACHE=/tmp/pvcache
  B="python bench.py --steps 5 --warmup 3 --shoot-photons 0 --maps-photons 0 --no-cpu-baseline"
  $B > gpurun_out/rpre_bench_default.log 2>&1
  PV_LIBPV=$V $B > gpurun_out/rpre_bench_variant.log 2>&1
  python tools/summ.py 'gpurun_out/rpre_bench_*.log'        # same checksum_L expected: the variant must be bit-identical
fi
# 2. PV_DEVICES in the drop-in binary: same image from one and from two devices (byte-identical where Li is
#    deterministic: homogeneous medium, one light), and the wall time of the frame's volume term
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  T=$(mktemp -d); cd $T
  /root/repo/baseline/_ref/pbrt_b200 --quiet /root/repo/tests/scenes/cornell_e2e.pbrt 2> one.err; mv cornell_e2e.pfm one.pfm
  PV_DEVICES=0,1 /root/repo/baseline/_ref/pbrt_b200 --quiet /root/repo/tests/scenes/cornell_e2e.pbrt 2> two.err; mv cornell_e2e.pfm two.pfm
  cmp one.pfm two.pfm && echo "PV_DEVICES=0,1: image identical to one device"
  grep "\[pv\]" one.err two.err
  cp one.err two.err /root/repo/gpurun_out/ 2>/dev/null
  cd /root/repo
fi
