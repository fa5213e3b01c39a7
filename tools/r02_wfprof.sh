# wavefront shooter: per-kernel time (ncu launch list) and one full capture of each of the three kernels
cd /root/repo
N=${1:-1000000}
python tools/shoot_probe.py $N > gpurun_out/wfp_plain.log 2>&1 || exit 1
tail -1 gpurun_out/wfp_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/wfp_launches.csv python tools/shoot_probe.py $N 1 > gpurun_out/wfp_ncu_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(l for l in open("gpurun_out/wfp_launches.csv") if l.startswith('"'))]
h = rows[0]; ik = h.index("Kernel Name"); iv = h.index("Metric Value")
t = collections.OrderedDict()
for r in rows[1:]:
    t.setdefault(r[ik].split("(")[0], []).append(float(r[iv].replace(",", "")))
for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])):
    print("%-60s n=%4d total %.3f ms avg %.4f ms" % (k[:60], len(v), sum(v) / 1e6, sum(v) / len(v) / 1e6))
PY
for k in wf_march_kernel wf_trace_kernel wf_event_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 30 -c 1 -o gpurun_out/wfp_$k -f python tools/shoot_probe.py $N 1 > gpurun_out/wfp_ncu_$k.log 2>&1
done
