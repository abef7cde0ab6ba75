#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c21
timeout -k 10 1500 python -m pytest tests -x -q -m gpu > $O.pytest.log 2>&1; echo "rc=$?" >> $O.pytest.log
tail -n 6 $O.pytest.log
timeout -k 10 600 python scripts/tv_gemm_perf.py > $O.perf.log 2>&1; echo "perf rc=$?"
grep -E '^(digits6)' $O.perf.log | cut -c1-330
