#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2san
timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest "tests/test_gemm_digits_gpu.py::test_gemm_digits_vs_numpy" "tests/test_gemm_digits_gpu.py::test_gemm_digits_alpha_beta_and_zero_rows" "tests/test_tv_plda_gpu.py::test_em_iteration" "tests/test_tv_plda_gpu.py::test_subtract_tett_ivectors" "tests/test_gmm_gpu.py::test_jfa_bwstats" -x -q -m gpu > $O.memcheck.log 2>&1; echo "memcheck rc=$?"
grep -c 'Invalid\|out of bounds' $O.memcheck.log; tail -n 6 $O.memcheck.log
timeout -k 10 600 python -m pytest tests/test_cli_gpu.py -x -q -m gpu 2>&1 | tail -3
