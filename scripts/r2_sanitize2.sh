#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2san2
timeout -k 10 100 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest "tests/test_tv_plda_gpu.py::test_em_iteration" "tests/test_tv_plda_gpu.py::test_subtract_tett_ivectors" "tests/test_gmm_gpu.py::test_jfa_normalize_features[256c-simt]" -x -q -m gpu > $O.memcheck.log 2>&1; echo "memcheck rc=$?"
grep -c 'Invalid\|out of bounds' $O.memcheck.log; tail -n 5 $O.memcheck.log | cut -c1-200
