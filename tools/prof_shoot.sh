cd /root/repo
ncu --set full --clock-control none --import-source on -k shoot_kernel --launch-skip 0 -c 1 -o gpurun_out/prof_shoot2 -f python tools/shoot_probe.py > gpurun_out/ncu_shoot2.log 2>&1
