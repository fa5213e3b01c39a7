cd /root/repo
python tools/shoot_probe.py > gpurun_out/shoot_probe.log 2>&1
for v in "$@"; do PV_LIBPV=/root/repo/cs348b-pbrt_b200/csrc/variants/libpv_$v.so python tools/shoot_probe.py >> gpurun_out/shoot_probe.log 2>&1; done
