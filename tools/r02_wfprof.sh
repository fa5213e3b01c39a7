# wavefront shooter: per-kernel time (ncu launch list) and one full capture of each of the three kernels in a REPRESENTATIVE
# generation (the first wave of pv_shoot is 64 blocks = a few small generations; launch 12 of each kernel is an early generation
# of the main wave, all slots busy)
cd /root/repo
N=${1:-2000000}; tag=${2:-r02_v1b}
python tools/shoot_probe.py $N > gpurun_out/${tag}_shoot_plain.log 2>&1 || exit 1
tail -1 gpurun_out/${tag}_shoot_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches.csv python tools/shoot_probe.py $N 1 > gpurun_out/${tag}_ncu_launches.log 2>&1
for k in wf_march_kernel wf_trace_kernel wf_event_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 11 -c 1 -o gpurun_out/${tag}_$k -f python tools/shoot_probe.py $N 1 > gpurun_out/${tag}_ncu_$k.log 2>&1
done
