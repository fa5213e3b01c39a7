# usage: bash tools/prof_gather.sh <tag> [PV_LIBPV path]
set -x
cd /root/repo
tag=$1
[ -n "$2" ] && export PV_LIBPV=$2
export PV_BENCH_CACHE=/tmp/pvcache
python bench.py --steps 1 --shoot-photons 0 --no-cpu-baseline > gpurun_out/plain_$tag.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k gather_kernel --launch-skip 3 -c 1 -o gpurun_out/prof_$tag -f \
    python bench.py --steps 1 --shoot-photons 0 --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1
