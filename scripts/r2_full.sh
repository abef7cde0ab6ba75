#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2full
timeout -k 10 1500 python -m pytest tests -x -q -m gpu > $O.pytest.log 2>&1; echo "rc=$?" >> $O.pytest.log
tail -n 5 $O.pytest.log
timeout -k 10 600 python bench.py --steps 10 --warmup 3 > $O.bench.json 2> $O.bench.err; echo "rc=$?"
cut -c1-1500 $O.bench.json
