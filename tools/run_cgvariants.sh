#!/bin/bash
# A/B of cellgather variants on the GPU box: usage tools/run_cgvariants.sh name [name ...]   (built by csrc/cgvariants.sh)
cd /root/repo
export PV_BENCH_CACHE=/tmp/pvcache
B="python bench.py --steps 4 --warmup 3 --shoot-photons 0 --no-cpu-baseline"
summ() { python -c "
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d['roofline']
    print(sys.argv[2], 'frame %.2f ms' % d['ms_per_step'], {k: round(v,2) for k,v in r['phase_ms'].items()}, 'march %.2f' % r['march_kernels_ms'], 'cand %.1f' % r['candidates_per_lookup'], d['checksum_L'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
" $1 $2; }
$B > gpurun_out/v_default.json 2> gpurun_out/v_default.err; summ gpurun_out/v_default.json default
for v in "$@"; do
  PV_LIBPV=/root/repo/cs348b-pbrt_b200/csrc/variants/libpv_$v.so $B > gpurun_out/v_$v.json 2> gpurun_out/v_$v.err; summ gpurun_out/v_$v.json $v
done
