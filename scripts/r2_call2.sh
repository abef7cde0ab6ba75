#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c2
timeout -k 10 300 python -m pytest tests/test_gmm_gpu.py -x -q -m gpu -k "tc and not tc2p" > $O.pytest_gmm.log 2>&1; echo "rc=$?" >> $O.pytest_gmm.log
timeout -k 10 300 python bench.py --kernel 2 --frames 4000000 --steps 5 --no-cpu-baseline --no-ivectors --e2e-steps 1 > $O.bench_k2.log 2>&1; echo "rc=$?" >> $O.bench_k2.log
LR_TC_PROF=1 timeout -k 10 300 python bench.py --kernel 2 --frames 4000000 --steps 2 --no-cpu-baseline --no-ivectors --e2e-steps 1 > $O.prof_k2.log 2>&1; echo "rc=$?" >> $O.prof_k2.log
timeout -k 10 300 python bench.py --kernel 3 --frames 4000000 --steps 5 --no-cpu-baseline --no-ivectors --e2e-steps 1 > $O.bench_k3.log 2>&1; echo "rc=$?" >> $O.bench_k3.log
tail -n 4 $O.pytest_gmm.log
grep -h tc_prof $O.prof_k2.log | tail -n 4
