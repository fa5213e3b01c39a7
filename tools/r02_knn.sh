cd /root/repo
python -m pytest tests/test_gpu_parity.py tests/test_gpu_config2.py tests/test_gpu_volint.py -m gpu -x -q > gpurun_out/knn_tests.log 2>&1; tail -5 gpurun_out/knn_tests.log
export PV_BENCH_CACHE=/tmp/pvcache
B="python bench.py --workload config2 --steps 5 --warmup 3 --shoot-photons 0 --maps-photons 0 --no-cpu-baseline"
$B > gpurun_out/knn_new.json 2> gpurun_out/knn_new.err; python tools/summ.py gpurun_out/knn_new.json
PV_KNN_LEGACY=1 $B > gpurun_out/knn_old.json 2> gpurun_out/knn_old.err; python tools/summ.py gpurun_out/knn_old.json
