#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2final4
timeout -k 10 1500 python -m pytest tests -q -m gpu > $O.pytest.log 2>&1; echo "rc=$?" >> $O.pytest.log
tail -n 4 $O.pytest.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout -k 10 900 python bench.py > $O.bench.json 2> $O.bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2final4.bench.json") if l.startswith("{")][-1])
r=d["roofline"]
print("value", d["value"], "ms", d["ms_per_step"], "frac", r["frac"], "kernel", r["kernel_frac"], "ceiling", r["ceiling"], "e2e", d["e2e"]["value"], d["clocks"])
for k in ("ivectors","tv_em","cpu_baseline"):
    print(k, json.dumps(d.get(k))[:200])
PY
