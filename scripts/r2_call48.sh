#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c48
timeout -k 10 900 python -m pytest tests/test_cli_gpu.py -x -q -m gpu -k "window_llr or ivector_mode" > $O.pytest.log 2>&1; echo "rc=$?" >> $O.pytest.log
tail -n 30 $O.pytest.log | cut -c1-300
