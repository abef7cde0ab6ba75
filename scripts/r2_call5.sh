#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 90 python -m pytest tests/test_gmm_gpu.py -x -q -m gpu -k "tc and not tc2p" 2>&1 | tail -n 3
for k in 2 3 2; do
timeout -k 5 60 python bench.py --kernel $k --frames 4000000 --steps 3 --warmup 3 --no-cpu-baseline --no-ivectors --e2e-steps 1 2>/dev/null | cut -c1-200
done
O=gpurun_out/r2ncu
timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:k_tc_one -s 3 -c 1 -o $O \
  python bench.py --kernel 2 --frames 2000000 --steps 1 --warmup 3 --no-cpu-baseline --no-ivectors --e2e-steps 1 > $O.log 2>&1
