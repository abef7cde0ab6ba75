#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c16
timeout -k 10 300 python -m pytest tests/test_gemm_digits_gpu.py -x -q -m gpu > $O.pytest_digits.log 2>&1; echo "rc=$?" >> $O.pytest_digits.log
tail -n 15 $O.pytest_digits.log
timeout -k 10 600 python -m pytest tests/test_tv_plda_gpu.py -x -q -m gpu > $O.pytest_tv.log 2>&1; echo "rc=$?" >> $O.pytest_tv.log
tail -n 4 $O.pytest_tv.log
timeout -k 10 600 python scripts/tv_gemm_perf.py > $O.perf.log 2>&1; echo "perf rc=$?"
grep -E '^(cublas|digits)' $O.perf.log | cut -c1-330
LR_I8_CLUSTER=1 timeout -k 10 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_gemm_i8 --csv --log-file $O.probe.csv python scripts/i8_probe.py > $O.probe.log 2>&1
grep k_gemm_i8 $O.probe.csv | awk -F'","' '{print $(NF-2), $(NF)}' | tail -2
