#!/bin/bash
# compute-sanitizer over the smoke case (shoot 2 k photons -> build the grid -> gather 96 rays -> check against the oracle):
# memcheck, racecheck (shared-memory hazards of the radix sort, the TMA-staged cell gather, the warp-shared lists) and synccheck.
# usage (GPU box): bash tools/sanitize.sh [tag]    -> gpurun_out/<tag>_{memcheck,racecheck,synccheck}.log + <tag>_summary.txt
cd /root/repo
tag=${1:-r02_sanitize}
python __graft_entry__.py smoke > gpurun_out/${tag}_plain.log 2>&1 || { echo "smoke failed without the sanitizer"; exit 1; }
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/${tag}_$tool.log --print-limit 20 python __graft_entry__.py smoke > gpurun_out/${tag}_$tool.out 2>&1
  echo "$tool: exit $? | $(tail -1 gpurun_out/${tag}_$tool.out) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${tag}_$tool.log | tail -1)"
done | tee gpurun_out/${tag}_summary.txt
# the k-nearest gather kernel and the all-maps shooter are not on the smoke path: one parity test each under memcheck
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/${tag}_memcheck_tests.log --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "knn_bit_exact_vs_reference and cornell_homog or all_maps_vs_oracle and cornell_surf or deep_continuation" > gpurun_out/${tag}_memcheck_tests.out 2>&1
echo "memcheck over parity tests (k-nearest lookup, all-maps shooter, deep stacks): exit $? | $(tail -1 gpurun_out/${tag}_memcheck_tests.out) | $(grep 'ERROR SUMMARY' gpurun_out/${tag}_memcheck_tests.log | tail -1)" | tee -a gpurun_out/${tag}_summary.txt
