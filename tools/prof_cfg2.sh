cd /root/repo
ncu --set full --clock-control none --import-source on -k gather_kernel --launch-skip 3 -c 1 -o gpurun_out/prof_cfg2 -f python bench.py --workload config2 --steps 1 --shoot-photons 0 --no-cpu-baseline > gpurun_out/ncu_cfg2.log 2>&1
