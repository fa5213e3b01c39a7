#!/bin/bash
# Round-2 evidence on ONE B200: the default bench (both arms), smoke, then -- each only after the same command has exited 0 without
# ncu -- the launch list and one full capture per kernel of the path; the sanitizer runs last.
# usage (GPU box): bash tools/r02_final_1gpu.sh [tag]      outputs under gpurun_out/<tag>_*
cd /root/repo
tag=${1:-r02_v1}
( time python bench.py ) > gpurun_out/${tag}_bench.log 2>&1; tail -1 gpurun_out/${tag}_bench.log | cut -c1-300
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/${tag}_ref.log 2>&1; tail -1 gpurun_out/${tag}_ref.log | cut -c1-300
( time python __graft_entry__.py smoke ) > gpurun_out/${tag}_smoke.log 2>&1; tail -4 gpurun_out/${tag}_smoke.log | head -1
export PV_BENCH_CACHE=/tmp/pvcache
B="python bench.py --steps 2 --warmup 3 --shoot-photons 2000000 --no-cpu-baseline"
$B > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_launches.log 2>&1
cap() { # kernel regex, label, launch-skip, command...
  k=$1; label=$2; skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip $skip -c 1 -o gpurun_out/${tag}_$label -f "$@" > gpurun_out/${tag}_ncu_$label.log 2>&1
  echo "$label: $(grep -c 'Profiling' gpurun_out/${tag}_ncu_$label.log) capture(s)"
}
G="python bench.py --steps 1 --warmup 3 --shoot-photons 0 --no-cpu-baseline"
cap cellgather_kernel cellgather 3 $G
cap recurrence_thread_kernel recurrence 3 $G
cap march_steps_kernel march 3 $G
cap rs_scatter_kernel rs_scatter 4 $G
S="python tools/shoot_probe.py 2000000 1"
$S > gpurun_out/${tag}_shoot_plain.log 2>&1
cap wf_march_kernel wf_march 30 $S
cap wf_event_kernel wf_event 30 $S
cap wf_trace_kernel wf_trace 30 $S
K="python bench.py --workload config2 --steps 1 --warmup 3 --shoot-photons 0 --maps-photons 0 --no-cpu-baseline"
$K > gpurun_out/${tag}_config2_plain.log 2>&1
cap cellgather_kernel cellknn 3 $K
python bench.py --workload config2 --steps 5 --warmup 3 --shoot-photons 0 --maps-photons 0 --no-cpu-baseline > gpurun_out/${tag}_config2.log 2>&1
bash tools/sanitize.sh ${tag}_sanitize
