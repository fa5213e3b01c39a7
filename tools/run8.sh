cd /root/repo
python -m pytest tests -x -q -m gpu -k "shooter or smoke or dropin" > gpurun_out/t13.log 2>&1; tail -3 gpurun_out/t13.log
export PV_BENCH_CACHE=/tmp/pvcache
python bench.py --steps 2 --no-cpu-baseline > gpurun_out/b13_default.log 2>&1
