#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c14
for dbg in 0 1 2; do
LR_I8_CLUSTER=1 LR_I8_DBG=$dbg timeout -k 10 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_gemm_i8 --csv --log-file $O.dbg$dbg.csv python scripts/i8_probe.py > $O.dbg$dbg.log 2>&1
echo "dbg $dbg rc=$?"; grep k_gemm_i8 $O.dbg$dbg.csv | awk -F'","' '{print $(NF-2), $(NF)}' | tail -4
done
