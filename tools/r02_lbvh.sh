#!/bin/bash
# round 2: LBVH + EXR tests first (new), then the whole GPU suite
python -m pytest tests/test_gpu_lbvh.py -m gpu -x -q -s 2>&1 | tail -15 > gpurun_out/lbvh_tests.log
python -m pytest tests/test_dropin_render.py -m gpu -q -k "exr or bvh" 2>&1 | tail -15 >> gpurun_out/lbvh_tests.log
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/all_tests.log
cat gpurun_out/lbvh_tests.log gpurun_out/all_tests.log
