cd /root/repo
python -m pytest tests/test_gpu_parity.py tests/test_gpu_config2.py tests/test_gpu_fullsize.py tests/test_gpu_volint.py -m gpu -x -q > gpurun_out/hist_tests.log 2>&1; tail -4 gpurun_out/hist_tests.log
export PV_BENCH_CACHE=/tmp/pvcache
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --maps-photons 0 > gpurun_out/hist_bench.log 2>&1
PV_KNN_NOHIST=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --maps-photons 0 > gpurun_out/nohist_bench.log 2>&1
python - <<'PY'
import json
for f in ("hist", "nohist"):
    try:
        l = [x for x in open("gpurun_out/%s_bench.log" % f) if x.startswith('{"metric"')][-1]; d = json.loads(l)
        fr = d["frame"]
        print(f, "step %.2f ms" % d["ms_per_step"], "frame gather %.1f ms" % fr["gather_device_ms"], {k: round(v, 2) for k, v in fr["gather_phase_ms_rank0"].items()}, fr["gather_L_sum"], d["checksum_L"])
    except Exception as e:
        print(f, "FAILED", e)
PY
