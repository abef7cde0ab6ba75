#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2ncu_chol
U3=1024 U4=128 timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:k_chol_fused -s 1 -c 1 -o $O \
  python scripts/tv_breakdown.py > $O.log 2>&1
echo "rc=$?" >> $O.log
tail -n 3 $O.log
