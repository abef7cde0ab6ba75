#!/bin/bash
# round-2 GPU call 1: TMEM shape probe, the three round-1 prototypes, parity of the one-pass kernel, timing
mkdir -p gpurun_out
O=gpurun_out/r2c1
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O.smi.log 2>&1
timeout -k 5 60 ./scripts/tmem_shape_probe > $O.probe.log 2>&1; echo "probe rc=$?" >> $O.probe.log
timeout -k 5 60 ./scripts/xchg_proto > $O.xchg.log 2>&1; echo "rc=$?" >> $O.xchg.log
timeout -k 5 90 ./scripts/onepass_bw_proto > $O.onepass_proto.log 2>&1; echo "rc=$?" >> $O.onepass_proto.log
timeout -k 5 90 ./scripts/gemm_split_proto > $O.gemm_split.log 2>&1; echo "rc=$?" >> $O.gemm_split.log
timeout -k 10 900 python -m pytest tests/test_gmm_gpu.py -x -q -m gpu > $O.pytest_gmm.log 2>&1; echo "rc=$?" >> $O.pytest_gmm.log
for k in 2 3; do
  timeout -k 10 300 python bench.py --kernel $k --frames 4000000 --steps 5 --no-cpu-baseline --no-ivectors --e2e-steps 1 > $O.bench_k$k.log 2>&1; echo "rc=$?" >> $O.bench_k$k.log
done
tail -5 $O.probe.log $O.xchg.log $O.onepass_proto.log $O.gemm_split.log
tail -15 $O.pytest_gmm.log
tail -3 $O.bench_k2.log $O.bench_k3.log
