cd /root/repo
python -m pytest tests -x -q -m gpu -k "shooter or dropin" > gpurun_out/t14.log 2>&1; tail -3 gpurun_out/t14.log
python tools/shoot_probe.py > gpurun_out/shoot_probe.log 2>&1
for v in "$@"; do PV_LIBPV=/root/repo/cs348b-pbrt_b200/csrc/variants/libpv_$v.so python tools/shoot_probe.py >> gpurun_out/shoot_probe.log 2>&1; done
