# The reference project's scenes AT THEIR SHIPPED SIZES through the drop-in binary (timings only; the parity tests use reduced sizes).
# usage: tools/e2e_full.sh [ref]   -- with "ref" the unmodified reference (oracle/_ref/pbrt_ref, all host cores) is timed too
cd /root/repo
mkdir -p gpurun_out/e2e_full && cd gpurun_out/e2e_full
python - <<'PY'
import sys; sys.path.insert(0, "/root/repo")
from __graft_entry__ import load_package
load_package()
from cs348b_pbrt_b200 import scenes
open("scene.pbrt", "w").write(scenes.sphere_pbrt(outfile="scene.pfm"))                 # projectScene/scene.pbrt: 1 M volume + 50 k caustic photons, 300x300, 8 spp
open("pinkfloyd.pbrt", "w").write(scenes.prism_pbrt(outfile="pinkfloyd.pfm"))          # projectScene/pinkfloyd.pbrt: 5 M volume photons, 512x512, 32 spp
open("volumescene.pbrt", "w").write(scenes.volumescene_pbrt(outfile="volumescene.pfm"))  # projectScene/volumescene_png.pbrt as shipped
PY
for s in ${SCENES:-volumescene scene}; do       # pinkfloyd (5 M photons, 512x512 x 32 spp, prism in view) takes minutes: SCENES=pinkfloyd
  t0=$(date +%s.%N); /root/repo/baseline/_ref/pbrt_b200 --quiet $s.pbrt 2>&1 | grep "\[pv\]"; t1=$(date +%s.%N)
  echo "$s drop-in wall $(python -c "print(round($t1 - $t0, 2))") s"
  if [ "$1" = ref ]; then t0=$(date +%s.%N); /root/repo/oracle/_ref/pbrt_ref --quiet --outfile ${s}_ref.pfm $s.pbrt > /dev/null 2>&1; t1=$(date +%s.%N)
    echo "$s reference wall $(python -c "print(round($t1 - $t0, 2))") s ($(nproc) cores)"; fi
done
