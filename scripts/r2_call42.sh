#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c42
timeout -k 10 300 python scripts/jfa_norm_perf.py 2>&1 | tail -1 | tee $O.jfa.log
U3=1024 U4=128 timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:k_chol_fused -s 1 -c 1 -o $O.chol \
  python scripts/tv_breakdown.py > $O.ncu.log 2>&1
echo "ncu rc=$?"
ls -la $O.chol.ncu-rep | cut -c1-120
