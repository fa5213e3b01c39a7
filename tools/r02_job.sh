cd /root/repo
T=$(mktemp -d); cd $T
for s in 0 1 2; do
  for n in volint_single_e2e volint_emission_e2e; do
    PV_SEED=$s /root/repo/baseline/_ref/pbrt_b200 --quiet /root/repo/tests/scenes/$n.pbrt 2> err_$n_$s.txt; cp $n.pfm /root/repo/gpurun_out/img_${n}_seed$s.pfm
  done
done
cd /root/repo
PV_TIMING=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_timing.log 2> gpurun_out/bench_timing.err
grep -c "pv timing" gpurun_out/bench_timing.err
