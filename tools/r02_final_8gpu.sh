#!/bin/bash
# Round-2 evidence on N GPUs of one box (gpurun --gpus N): the bench under torchrun at N (and the powers of two below it when
# SWEEP=1), the drop-in binary with PV_DEVICES over all N devices, BASELINE config 5 (128 M photons, 3840x2160) at N.
cd /root/repo
N=${1:-8}
for n in $( [ -n "$SWEEP" ] && echo "2 4 $N" | tr ' ' '\n' | awk -v N=$N '$1<=N' | sort -nu | tr '\n' ' ' || echo $N ); do
  if [ "$n" = 1 ]; then python bench.py --no-cpu-baseline > gpurun_out/r02_scale_n1.log 2>&1
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n > gpurun_out/r02_scale_n$n.log 2>&1; fi
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    l = [x for x in open("gpurun_out/r02_scale_n%s.log" % n) if x.startswith('{"metric"')][-1]; d = json.loads(l)
    f = d.get("frame") or {}
    print("N=%s: %.2f M rays/s (%.2f ms/step), e2e %.2f M rays/s; frame %.0f ms (shoot %.0f dev, all-gather %s, build %.1f, gather %.1f); allgather %s" % (
        n, d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, f.get("frame_wall_ms", 0), f.get("shoot_device_ms", 0),
        (f.get("allgather") or {}).get("collective_ms"), f.get("build_wall_ms", 0), f.get("gather_device_ms", 0), d.get("allgather")))
except Exception as e:
    print("N=%s FAILED" % n, e)
PY
done
T=$(mktemp -d); cd $T
DEVS=$(python -c "print(','.join(str(i) for i in range($N)))")
for s in cornell_e2e sphere_e2e; do
  /root/repo/baseline/_ref/pbrt_b200 --quiet /root/repo/tests/scenes/$s.pbrt 2> one_$s.err; mv $s.pfm one_$s.pfm
  PV_DEVICES=$DEVS /root/repo/baseline/_ref/pbrt_b200 --quiet /root/repo/tests/scenes/$s.pbrt 2> many_$s.err; mv $s.pfm many_$s.pfm
  cmp one_$s.pfm many_$s.pfm && echo "$s: PV_DEVICES=$DEVS image identical to one device"
  ( echo "== $s, one device"; grep "\[pv\]" one_$s.err; echo "== $s, PV_DEVICES=$DEVS"; grep "\[pv\]" many_$s.err ) >> /root/repo/gpurun_out/r02_pv_devices_n$N.log
done
cat /root/repo/gpurun_out/r02_pv_devices_n$N.log
cd /root/repo
[ -n "$CFG5" ] && { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload config5 --steps 3 --warmup 3 --shoot-photons 0 > gpurun_out/r02_cfg5_n$N.log 2>&1; tail -1 gpurun_out/r02_cfg5_n$N.log | cut -c1-600; }
