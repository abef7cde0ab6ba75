#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c34
timeout -k 10 900 python -m pytest tests/test_gmm_gpu.py tests/test_cli_gpu.py -x -q -m gpu -k "train_world" > $O.pytest.log 2>&1; echo "rc=$?" >> $O.pytest.log
tail -n 30 $O.pytest.log | cut -c1-300
