#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c4
timeout -k 5 90 python -m pytest tests/test_gmm_gpu.py -x -q -m gpu -k "tc and not tc2p" 2>&1 | tail -n 5
for d in 0; do
echo "== dbg $d"
LR_TC_DEBUG=$d LR_TC_PROF=1 timeout -k 5 60 python bench.py --kernel 2 --frames 4000000 --steps 1 --warmup 3 --no-cpu-baseline --no-ivectors --e2e-steps 1 > $O.prof$d.log 2>&1; echo "rc=$?" >> $O.prof$d.log
grep -h tc_prof $O.prof$d.log | tail -n 6
LR_TC_DEBUG=$d timeout -k 5 60 python bench.py --kernel 2 --frames 4000000 --steps 3 --warmup 3 --no-cpu-baseline --no-ivectors --e2e-steps 1 2>/dev/null | cut -c1-200
done
