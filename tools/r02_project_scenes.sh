#!/bin/bash
# Every .pbrt file the reference project ships (projectScene/, staged unmodified under baseline/_ref/projectScene by the builder:
# git-ignored, travels to the GPU box), rendered by the drop-in AT ITS SHIPPED SIZE on one B200.  Output: one block per scene with
# the adapter's phase lines and the wall time.
cd /root/repo/baseline/_ref/projectScene || exit 1
for f in volumescene_png closeup_png closeup scene_png pinkfloyd_png rainbow2_png rainbow_png scene pinkfloyd darkside; do
  echo "== $f.pbrt"
  grep -o '"integer [xy]resolution" \[[0-9]*\]\|"integer pixelsamples" \[[0-9]*\]\|"integer volumephotons" *\[[0-9]*\]\|"integer nused" \[[0-9]*\]' $f.pbrt | grep -v "851\|315" | tr '\n' ' '; echo
  ( time timeout 600 /root/repo/baseline/_ref/pbrt_b200 --quiet $f.pbrt ) 2>&1 | grep "^\[pv\]\|^real\|Error\|Severe" | grep -v ioctl | cut -c1-260
done
ls -la *.exr *.png 2>/dev/null | awk '{print $5, $9}'
