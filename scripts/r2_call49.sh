#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c49
timeout -k 10 900 python -m pytest tests/test_cli_gpu.py -x -q -m gpu -k "init_by_client" > $O.pytest.log 2>&1; echo "rc=$?" >> $O.pytest.log
tail -n 30 $O.pytest.log | cut -c1-300
