#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c11
timeout -k 10 600 python -m pytest tests/test_gemm_digits_gpu.py -x -q -m gpu > $O.pytest_digits.log 2>&1; echo "rc=$?" >> $O.pytest_digits.log
tail -n 25 $O.pytest_digits.log
timeout -k 10 600 python -m pytest tests/test_tv_plda_gpu.py -x -q -m gpu > $O.pytest_tv.log 2>&1; echo "rc=$?" >> $O.pytest_tv.log
tail -n 8 $O.pytest_tv.log
timeout -k 10 600 python scripts/tv_gemm_perf.py > $O.perf.log 2>&1; echo "perf rc=$?"
tail -n 12 $O.perf.log
