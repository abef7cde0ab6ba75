#!/bin/bash
timeout -k 5 90 python -m pytest tests/test_gmm_gpu.py -x -q -m gpu -k "tc and not tc2p" 2>&1 | tail -n 3
for d in 0; do
LR_TC_DEBUG=$d timeout -k 5 60 python bench.py --kernel 2 --frames 4000000 --steps 3 --warmup 3 --no-cpu-baseline --no-ivectors --e2e-steps 1 2>/dev/null | cut -c1-200
done
