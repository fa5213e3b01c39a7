cd /root/repo
export PV_BENCH_CACHE=/tmp/pvcache
B="python bench.py --steps 3 --shoot-photons 0 --no-cpu-baseline"
$B > gpurun_out/b16_default.log 2>&1
for v in "$@"; do
  PV_LIBPV=/root/repo/cs348b-pbrt_b200/csrc/variants/libpv_$v.so $B > gpurun_out/b16_$v.log 2>&1
done
