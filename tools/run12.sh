cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/t18.log 2>&1; tail -3 gpurun_out/t18.log
for c in 1.0 0.8 0.7 0.6 0.5; do
  PV_KNN_CELL=$c python bench.py --workload config2 --steps 3 --shoot-photons 0 --no-cpu-baseline > gpurun_out/b15_knn$c.log 2>&1
done
