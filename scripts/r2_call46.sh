#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c46
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O.jfa_launches.csv python scripts/jfa_norm_perf.py > $O.log 2>&1; echo "rc=$?"
python scripts/launch_summary.py $O.jfa_launches.csv --title jfa | tail -14
