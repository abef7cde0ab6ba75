#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c3
LR_TC_PROF=1 timeout -k 10 300 python bench.py --kernel 2 --frames 4000000 --steps 1 --warmup 3 --no-cpu-baseline --no-ivectors --e2e-steps 1 > $O.prof_k2.log 2>&1; echo "rc=$?" >> $O.prof_k2.log
grep -h tc_cta $O.prof_k2.log | tail -n 144 > $O.cta.log
grep -h tc_prof $O.prof_k2.log | tail -n 4
