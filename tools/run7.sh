cd /root/repo
export PV_BENCH_CACHE=/tmp/pvcache
B="python bench.py --steps 3 --shoot-photons 0 --no-cpu-baseline"
$B > gpurun_out/b12_default.log 2>&1
PV_LIBPV=/root/repo/cs348b-pbrt_b200/csrc/variants/libpv_lr0.so $B > gpurun_out/b12_lr0.log 2>&1
S="python bench.py --steps 1 --warmup 1 --shoot-photons 400000 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k shoot_kernel --launch-skip 1 -c 1 -o gpurun_out/prof_shoot -f $S > gpurun_out/ncu_shoot.log 2>&1
