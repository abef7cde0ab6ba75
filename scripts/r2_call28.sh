#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c28
timeout -k 10 600 python -m pytest tests/test_tv_plda_gpu.py tests/test_gemm_digits_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout -k 10 600 python scripts/tv_gemm_perf.py > $O.perf.log 2>&1; echo "perf rc=$?"
grep -E '^(digits6)' $O.perf.log | cut -c1-330
timeout -k 10 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O.bench.json 2> $O.bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2c28.bench.json") if l.startswith("{")][-1])
print("value", d["value"], "frac", d["roofline"]["frac"])
for k in ("ivectors","ivector_pipeline","tv_em"):
    print(k, json.dumps(d.get(k))[:260])
PY
