#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c13
timeout -k 10 900 python -m pytest tests/test_gemm_digits_gpu.py tests/test_tv_plda_gpu.py tests/test_gmm_gpu.py tests/test_cli_gpu.py -x -q -m gpu > $O.pytest.log 2>&1; echo "rc=$?" >> $O.pytest.log
tail -n 12 $O.pytest.log
for cs in 1 2 4; do
  LR_I8_CLUSTER=$cs timeout -k 10 300 python -m pytest tests/test_gemm_digits_gpu.py -x -q -m gpu > $O.pytest_c$cs.log 2>&1; echo "cluster $cs pytest rc=$?"
  LR_I8_CLUSTER=$cs timeout -k 10 600 python scripts/tv_gemm_perf.py > $O.perf_c$cs.log 2>&1; echo "cluster $cs perf rc=$?"
  grep -E '^(cublas|digits)' $O.perf_c$cs.log | cut -c1-330
done
