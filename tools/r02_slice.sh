cd /root/repo
export PV_BENCH_CACHE=/tmp/pvcache
B="python bench.py --slice-of ${SLICE:-8} --steps 5 --warmup 3 --shoot-photons 0 --maps-photons 0 --no-cpu-baseline --no-extra"
run() { label=$1; shift; env "$@" $B > gpurun_out/slice_$label.json 2> gpurun_out/slice_$label.err
  python - "$label" <<'PY'
import json, sys
f = "gpurun_out/slice_%s.json" % sys.argv[1]
try:
    l = [x for x in open(f) if x.startswith('{')][-1]; d = json.loads(l); r = d['roofline']
    print("%s: step %.3f ms" % (sys.argv[1], d['ms_per_step']), {k: round(v, 3) for k, v in r['phase_ms'].items()}, "march %.3f" % r['march_kernels_ms'], d['checksum_L'])
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(f.replace('.json', '.err')).read()[-400:])
PY
}
V=/root/repo/cs348b-pbrt_b200/csrc/variants
run default A=1
for v in $VARIANTS; do run $v PV_LIBPV=$V/libpv_$v.so; done
