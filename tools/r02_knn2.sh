cd /root/repo
python -m pytest tests/test_gpu_parity.py tests/test_gpu_config2.py tests/test_gpu_volint.py -m gpu -x -q > gpurun_out/knn_tests.log 2>&1; tail -3 gpurun_out/knn_tests.log
export PV_BENCH_CACHE=/tmp/pvcache
B="python bench.py --workload config2 --steps 5 --warmup 3 --shoot-photons 0 --maps-photons 0 --no-cpu-baseline"
for cfg in ${CFGS:-"1.0 1.2" "1.0 1.1" "1.0 1.35" "1.5 1.1" "1.5 1.2" "0.75 1.2"}; do
  set -- $cfg
  PV_KNN_CELL=$1 PV_KNN_TRIAL=$2 $B > gpurun_out/knn_c$1_t$2.json 2> gpurun_out/knn_c$1_t$2.err
  python - "$1" "$2" <<'PY'
import json, sys
f = "gpurun_out/knn_c%s_t%s.json" % (sys.argv[1], sys.argv[2])
try:
    l = [x for x in open(f) if x.startswith('{')][-1]; d = json.loads(l); r = d['roofline']
    print("cell %s trial %s: step %.2f ms" % (sys.argv[1], sys.argv[2], d['ms_per_step']), {k: round(v, 2) for k, v in r['phase_ms'].items()}, "cand %.0f" % r['candidates_per_lookup'], d['checksum_L'])
except Exception as e:
    print("cell %s trial %s FAILED" % (sys.argv[1], sys.argv[2]), e, open(f.replace('.json', '.err')).read()[-500:])
PY
done
