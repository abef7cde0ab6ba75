#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c10
timeout -k 10 1500 python -m pytest tests -x -q -m gpu > $O.pytest.log 2>&1; echo "rc=$?" >> $O.pytest.log
tail -n 5 $O.pytest.log
timeout -k 10 600 python bench.py --steps 10 --warmup 3 > $O.bench.json 2> $O.bench.err; echo "rc=$?"
LR_TC_NOSTAGGER=1 timeout -k 10 600 python bench.py --steps 10 --warmup 3 > $O.bench_nostagger.json 2> $O.bench_nostagger.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2c10.bench.json","gpurun_out/r2c10.bench_nostagger.json"):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["clocks"])
    except Exception as ex: print(f, "ERR", ex)
PY
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O.launches.csv python bench.py --steps 2 --warmup 1 > $O.ncu_bench.log 2>&1; echo "ncu rc=$?"
