#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_tv_plda_gpu.py -x -q -m gpu -k "plda" 2>&1 | tail -n 3
timeout -k 10 200 python scripts/plda_perf.py 2>&1 | tail -n 1
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_plda_launches.csv python scripts/plda_perf.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2_plda_launches.csv')) if len(r) > 10 and r[0].isdigit()]
agg = collections.Counter(); cnt = collections.Counter()
for r in rows:
    name = r[4][:60]; agg[name] += float(r[-1]); cnt[name] += 1
tot = sum(agg.values())
print("total kernel time (3 calls) %.2f ms" % (tot/1e6 if tot > 1e5 else tot))
for k, v in agg.most_common(12): print("%-62s n=%4d %12.1f" % (k, cnt[k], v))
PY
