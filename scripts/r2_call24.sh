#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c24
timeout -k 10 900 python -m pytest tests/test_multi_gpu.py tests/test_multi_gpu_cli.py -q -m gpu > $O.pytest2.log 2>&1; echo "rc=$?" >> $O.pytest2.log
tail -n 5 $O.pytest2.log
