#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2final5_2gpu
timeout -k 10 900 python -m pytest tests/test_multi_gpu.py tests/test_multi_gpu_cli.py -q -m gpu > $O.pytest.log 2>&1; echo "rc=$?" >> $O.pytest.log
tail -n 4 $O.pytest.log
