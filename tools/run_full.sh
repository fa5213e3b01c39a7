# full default bench + reference arm + smoke + ncu evidence (launch list, full capture of gather_kernel and march_steps_kernel)
cd /root/repo
tag=${1:-r01_v2}
( time python bench.py ) > gpurun_out/${tag}_bench.log 2>&1
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/${tag}_ref.log 2>&1
( time python __graft_entry__.py smoke ) > gpurun_out/${tag}_smoke.log 2>&1
export PV_BENCH_CACHE=/tmp/pvcache
B="python bench.py --steps 2 --warmup 3 --shoot-photons 100000 --no-cpu-baseline"
$B > gpurun_out/${tag}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k gather_kernel --launch-skip 3 -c 1 -o gpurun_out/${tag}_gather -f $B > gpurun_out/${tag}_ncu_gather.log 2>&1
ncu --set full --clock-control none --import-source on -k march_steps_kernel --launch-skip 3 -c 1 -o gpurun_out/${tag}_march -f $B > gpurun_out/${tag}_ncu_march.log 2>&1
# the shooter: volume-only and all-maps instantiations (the first launches of a run are the volume-only waves, then pv_shoot_maps)
ncu --set full --clock-control none --import-source on -k regex:shoot_kernel -c 10 -o gpurun_out/${tag}_shoot -f $B > gpurun_out/${tag}_ncu_shoot.log 2>&1
# BASELINE config 2 (k-nearest gather, 1 M photons, 512x512, k = 50) as a second bench line
python bench.py --workload config2 --steps 5 --warmup 3 --shoot-photons 0 --maps-photons 0 --no-cpu-baseline > gpurun_out/${tag}_config2.log 2>&1
# VolumeIntegrator "single" / "emission" (pv_volume_li) on the config-3 scene shape: plain run, then one full capture of the
# thread-per-ray recurrence kernel (the first timed "single" launch)
python tools/volint_bench.py > gpurun_out/${tag}_volint_bench.jsonl 2> gpurun_out/${tag}_volint_bench.err
ncu --set full --clock-control none --import-source on -k regex:volint_thread_kernel --launch-skip 3 -c 1 -o gpurun_out/${tag}_volint -f python tools/volint_bench.py --no-cpu --steps 1 --warmup 3 > gpurun_out/${tag}_ncu_volint.log 2>&1
python tools/volint_bench.py --impl reference > gpurun_out/${tag}_volint_ref.jsonl 2> gpurun_out/${tag}_volint_ref.err     # the reference's own classes on the box's host cores
