#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c44
timeout -k 10 600 python -m pytest tests/test_tv_plda_gpu.py tests/test_gemm_digits_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout -k 10 600 python scripts/tv_gemm_perf.py > $O.perf.log 2>&1; echo "perf rc=$?"
grep -E '^(digits6)' $O.perf.log | cut -c1-330
