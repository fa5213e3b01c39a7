cd /root/repo
export PV_BENCH_CACHE=/tmp/pvcache
B="python bench.py --workload config2 --steps 5 --warmup 3 --shoot-photons 0 --maps-photons 0 --no-cpu-baseline"
run() { # label, env...
  label=$1; shift
  env "$@" $B > gpurun_out/knn_$label.json 2> gpurun_out/knn_$label.err
  python - "$label" <<'PY'
import json, sys
f = "gpurun_out/knn_%s.json" % sys.argv[1]
try:
    l = [x for x in open(f) if x.startswith('{')][-1]; d = json.loads(l); r = d['roofline']
    print("%s: step %.2f ms" % (sys.argv[1], d['ms_per_step']), {k: round(v, 2) for k, v in r['phase_ms'].items()}, "cand %.0f" % r['candidates_per_lookup'], d['checksum_L'])
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(f.replace('.json', '.err')).read()[-500:])
PY
}
V=/root/repo/cs348b-pbrt_b200/csrc/variants
run t1.05 PV_KNN_TRIAL=1.05
run t1.0 PV_KNN_TRIAL=1.0
run t1.15 PV_KNN_TRIAL=1.15
run k80_t1.1 PV_KNN_TRIAL=1.1 PV_LIBPV=$V/libpv_k80.so
run k72_t1.1 PV_KNN_TRIAL=1.1 PV_LIBPV=$V/libpv_k72.so
run k96w4_t1.1 PV_KNN_TRIAL=1.1 PV_LIBPV=$V/libpv_k96w4.so
run c0.75_t1.1 PV_KNN_TRIAL=1.1 PV_KNN_CELL=0.75
run c0.75_t1.05 PV_KNN_TRIAL=1.05 PV_KNN_CELL=0.75
