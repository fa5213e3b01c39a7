set -x
cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/t6.log 2>&1; tail -5 gpurun_out/t6.log
export PV_BENCH_CACHE=/tmp/pvcache
python bench.py --steps 2 --shoot-photons 0 --no-cpu-baseline > gpurun_out/b6_default.log 2>&1
for v in s256w4c5 s256w2c11 s256w2c10 s512w2c8; do
  PV_LIBPV=/root/repo/cs348b-pbrt_b200/csrc/variants/libpv_$v.so python bench.py --steps 2 --shoot-photons 0 --no-cpu-baseline > gpurun_out/b6_$v.log 2>&1
done
grep -h -o '"value": [0-9.]*\|"frac": [0-9.]*\|"march_kernels_ms": [0-9.]*\|"avg_launch_ms": [0-9.]*\|checksum_L": [0-9.]*' gpurun_out/b6_*.log
