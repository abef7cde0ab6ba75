#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c25
for cs in 2 1; do
  LR_I8_CLUSTER=$cs timeout -k 10 600 python -m pytest tests/test_gemm_digits_gpu.py tests/test_tv_plda_gpu.py -x -q -m gpu > $O.pytest_c$cs.log 2>&1; echo "cluster $cs pytest rc=$?"; tail -n 2 $O.pytest_c$cs.log
  LR_I8_CLUSTER=$cs timeout -k 10 600 python scripts/tv_gemm_perf.py > $O.perf_c$cs.log 2>&1; echo "cluster $cs perf rc=$?"
  grep -E '^(digits6)' $O.perf_c$cs.log | cut -c1-330
  LR_I8_CLUSTER=$cs timeout -k 10 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_gemm_i8 --csv --log-file $O.probe_c$cs.csv python scripts/i8_probe.py > $O.probe_c$cs.log 2>&1
  grep k_gemm_i8 $O.probe_c$cs.csv | awk -F'","' '{print $(NF-2), $(NF)}' | tail -2
done
