#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c37
LR_CHOL_PROF=1 timeout -k 10 600 python scripts/tv_gemm_perf.py > $O.perf.log 2> $O.perf.err; echo "perf rc=$?"
grep chol_prof $O.perf.err | sort | uniq -c | sort -rn | head -8 | cut -c1-400
