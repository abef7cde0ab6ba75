#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2ncu
timeout -k 5 280 ncu --set full --clock-control none --import-source on -k regex:k_tc_one -s 3 -c 1 -o $O \
  python bench.py --kernel 2 --frames 2000000 --steps 1 --warmup 3 --no-cpu-baseline --no-ivectors --e2e-steps 1 > $O.log 2>&1
echo "rc=$?" >> $O.log
tail -n 5 $O.log
ls -la gpurun_out/ | tail -n 5
