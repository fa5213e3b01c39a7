#!/bin/bash
# pinkfloyd at its shipped size, 1 spp: the drop-in's phases under different photon-grid cell sizes (tuning probe)
cd /tmp
for c in ${CELLS:-default 0.5 0.25}; do
  if [ "$c" = default ]; then unset PV_KNN_CELL; else export PV_KNN_CELL=$c; fi
  echo "== PV_KNN_CELL=$c"
  ( time /root/repo/baseline/_ref/pbrt_b200 --quiet /root/repo/tests/scenes/pinkfloyd_1spp.pbrt ) 2>&1 | grep "^\[pv\]\|^real" | cut -c1-230
done
