#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c9
timeout -k 10 900 python -m pytest tests/test_multi_gpu_cli.py -q -m gpu > $O.pytest.log 2>&1; echo "rc=$?" >> $O.pytest.log
tail -n 30 $O.pytest.log
