#!/bin/bash
# Round-2 final evidence on ONE B200 (tag r02_v4): the whole GPU suite, the default bench (both arms), smoke, then -- each only after
# the same command has exited 0 without ncu -- the launch list of the bench and full captures of the LBVH build kernels.
cd /root/repo
tag=${1:-r02_v4}
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_tests.log
( time python bench.py ) > gpurun_out/${tag}_bench.log 2>&1; tail -4 gpurun_out/${tag}_bench.log | grep '^{' | cut -c1-200
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/${tag}_ref.log 2>&1; grep '^{' gpurun_out/${tag}_ref.log | cut -c1-200
( time python __graft_entry__.py smoke ) > gpurun_out/${tag}_smoke.log 2>&1; tail -4 gpurun_out/${tag}_smoke.log | head -1
export PV_BENCH_CACHE=/tmp/pvcache
B="python bench.py --steps 2 --warmup 3 --shoot-photons 2000000 --no-cpu-baseline"
$B > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_launches.log 2>&1
L="python tools/lbvh_probe.py 2000000 2"
$L > gpurun_out/${tag}_lbvh_plain.log 2>&1 || { echo "lbvh probe failed"; tail -5 gpurun_out/${tag}_lbvh_plain.log; exit 1; }
cat gpurun_out/${tag}_lbvh_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_lbvh_launches.csv $L > gpurun_out/${tag}_ncu_lbvh_launches.log 2>&1
for k in lbvh_refit_kernel lbvh_emit_kernel lbvh_hierarchy_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 1 -c 1 -o gpurun_out/${tag}_$k -f $L > gpurun_out/${tag}_ncu_$k.log 2>&1
  echo "$k: $(grep -c 'Profiling' gpurun_out/${tag}_ncu_$k.log) capture(s)"
done
