cd /root/repo
mkdir -p /tmp/e2e && cd /tmp/e2e
for n in config4_prism config1_volumescene; do
  /root/repo/baseline/_ref/pbrt_b200 --quiet /root/repo/tests/scenes/$n.pbrt > /root/repo/gpurun_out/e2e_$n.log 2>&1
  cp $n.pfm /root/repo/gpurun_out/e2e_$n.pfm
done
