# wavefront shooter: parity tests (shooter + all-maps + sharding), then A/B of variants / slot counts on the config-3 scene
cd /root/repo
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shooter or all_maps or radiance or smoke or bvh" > gpurun_out/wf_tests.log 2>&1; tail -3 gpurun_out/wf_tests.log
timeout 300 python tools/shoot_probe.py 2000000 > gpurun_out/wf_probe.log 2>&1; tail -1 gpurun_out/wf_probe.log
[ -n "$MEGA" ] && { PV_SHOOT_MEGAKERNEL=1 timeout 300 python tools/shoot_probe.py 2000000 > gpurun_out/mega_probe.log 2>&1; tail -1 gpurun_out/mega_probe.log; }
for v in ${VARIANTS}; do PV_WF_SLOTS=${VSLOTS:-524288} PV_LIBPV=/root/repo/cs348b-pbrt_b200/csrc/variants/libpv_$v.so timeout 300 python tools/shoot_probe.py 2000000 2 > gpurun_out/wf_probe_$v.log 2>&1; echo "$v: $(tail -1 gpurun_out/wf_probe_$v.log)"; done
for p in ${SLOTS}; do PV_WF_SLOTS=$p timeout 300 python tools/shoot_probe.py 2000000 2 > gpurun_out/wf_probe_$p.log 2>&1; echo "slots $p: $(tail -1 gpurun_out/wf_probe_$p.log)"; done
