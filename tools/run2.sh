set -x
cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/t7.log 2>&1; tail -5 gpurun_out/t7.log
export PV_BENCH_CACHE=/tmp/pvcache
B="python bench.py --steps 2 --shoot-photons 0 --no-cpu-baseline"
$B > gpurun_out/b7_default.log 2>&1
PV_XSHIFT_MAX=0 $B > gpurun_out/b7_xs0.log 2>&1
PV_XSHIFT_MAX=1 $B > gpurun_out/b7_xs1.log 2>&1
PV_XSHIFT_MAX=3 $B > gpurun_out/b7_xs3.log 2>&1
for v in s256c4 s256c5 m6 m8; do
  PV_LIBPV=/root/repo/cs348b-pbrt_b200/csrc/variants/libpv_$v.so $B > gpurun_out/b7_$v.log 2>&1
done
