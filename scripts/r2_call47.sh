#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c47
timeout -k 10 600 python -m pytest tests/test_gmm_gpu.py tests/test_cli_gpu.py -x -q -m gpu -k "jfa" 2>&1 | tail -3
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O.jfa_launches.csv python scripts/jfa_norm_perf.py > $O.log 2>&1; echo "rc=$?"; tail -1 $O.log | cut -c1-200
python scripts/launch_summary.py $O.jfa_launches.csv --title jfa | grep -E "k_jfa|k_tc_lse|Total"
timeout -k 10 300 python scripts/jfa_norm_perf.py 2>&1 | tail -1
