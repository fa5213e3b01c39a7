cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/t15.log 2>&1; tail -3 gpurun_out/t15.log
python tools/shoot_probe.py > gpurun_out/shoot_probe.log 2>&1
for v in "$@"; do PV_LIBPV=/root/repo/cs348b-pbrt_b200/csrc/variants/libpv_$v.so python tools/shoot_probe.py >> gpurun_out/shoot_probe.log 2>&1; done
export PV_BENCH_CACHE=/tmp/pvcache
B="python bench.py --steps 2 --shoot-photons 0 --no-cpu-baseline"
$B > gpurun_out/b14_default.log 2>&1
for v in "$@"; do PV_LIBPV=/root/repo/cs348b-pbrt_b200/csrc/variants/libpv_$v.so $B > gpurun_out/b14_$v.log 2>&1; done
