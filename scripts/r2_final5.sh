#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2final5
timeout -k 10 1500 python -m pytest tests -q -m gpu > $O.pytest.log 2>&1; echo "rc=$?" >> $O.pytest.log
tail -n 4 $O.pytest.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout -k 10 900 python bench.py > $O.bench.json 2> $O.bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2final5.bench.json") if l.startswith("{")][-1])
r=d["roofline"]
print("value", d["value"], "ms", d["ms_per_step"], "frac", r["frac"], "kernel", r["kernel_frac"], "e2e", d["e2e"]["value"], d["clocks"])
for k in ("ivectors","ivector_pipeline","tv_em","plda","product_levels"):
    print(k, json.dumps(d.get(k))[:230])
PY
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > $O.bench_ref.json 2> $O.bench_ref.err; echo "ref rc=$?"; tail -c 400 $O.bench_ref.json
U3=1024 U4=640 timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O.tv_launches.csv python scripts/tv_breakdown.py > $O.tv_breakdown.log 2>&1; echo "ncu tv rc=$?"
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O.bench_launches.csv python bench.py --steps 2 --warmup 3 --frames 2000000 --no-cpu-baseline --no-extra --no-ivectors --e2e-steps 1 > $O.ncu_bench.log 2>&1; echo "ncu bench rc=$?"
