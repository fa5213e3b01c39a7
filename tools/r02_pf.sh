#!/bin/bash
# pinkfloyd at its shipped size, 1 spp: the drop-in's phases under different photon-grid cell sizes (tuning probe)
cd /tmp
sed 's/"integer pixelsamples" \[32\]/"integer pixelsamples" [1]/' /root/repo/baseline/_ref/projectScene/pinkfloyd.pbrt > /tmp/pinkfloyd_1spp.pbrt; cp -r /root/repo/baseline/_ref/projectScene/obj /tmp/ 2>/dev/null
for c in ${CELLS:-default 0.5 0.25}; do
  if [ "$c" = default ]; then unset PV_KNN_CELL; else export PV_KNN_CELL=$c; fi
  echo "== PV_KNN_CELL=$c"
  ( time /root/repo/baseline/_ref/pbrt_b200 --quiet /tmp/pinkfloyd_1spp.pbrt ) 2>&1 | grep "^\[pv\]\|^real" | cut -c1-230
done
