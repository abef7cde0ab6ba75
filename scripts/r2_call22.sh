#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c22
timeout -k 10 600 python -m pytest tests/test_tv_plda_gpu.py tests/test_gemm_digits_gpu.py -x -q -m gpu > $O.pytest.log 2>&1; echo "rc=$?" >> $O.pytest.log
tail -n 4 $O.pytest.log
timeout -k 10 600 python scripts/tv_gemm_perf.py > $O.perf.log 2>&1; echo "perf rc=$?"
grep -E '^(digits6)' $O.perf.log | cut -c1-330
U3=1024 U4=640 timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O.tv_launches.csv python scripts/tv_breakdown.py > $O.tv_breakdown.log 2>&1; echo "ncu rc=$?"
