cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/t10.log 2>&1; tail -3 gpurun_out/t10.log
export PV_BENCH_CACHE=/tmp/pvcache
B="python bench.py --steps 3 --shoot-photons 0 --no-cpu-baseline"
for s in 1 2 4 7; do PV_GATHER_SLICES=$s $B > gpurun_out/b10_slices$s.log 2>&1; done
