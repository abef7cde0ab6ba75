#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2ncu_i8
# second k_gemm_i8 launch of the run = L = N TETt at R = 400, 1024 utterances (first one is the warm-up call's)
U3=1024 U4=128 timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:k_gemm_i8 -s 2 -c 2 -o $O \
  python scripts/tv_breakdown.py > $O.log 2>&1
echo "rc=$?" >> $O.log
tail -n 5 $O.log
timeout -k 10 300 python -m pytest tests/test_tv_plda_gpu.py tests/test_gemm_digits_gpu.py -x -q -m gpu 2>&1 | tail -3
