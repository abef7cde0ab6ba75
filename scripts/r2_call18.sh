#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c18
timeout -k 10 900 python -m pytest tests/test_gmm_gpu.py tests/test_tv_plda_gpu.py -x -q -m gpu > $O.pytest.log 2>&1; echo "rc=$?" >> $O.pytest.log
tail -n 12 $O.pytest.log
timeout -k 10 600 python scripts/tv_gemm_perf.py > $O.perf.log 2>&1; echo "perf rc=$?"
grep -E '^(digits6)' $O.perf.log | cut -c1-330
timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-extra > $O.bench.json 2> $O.bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2c18.bench.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_frac"], d["e2e"], d["clocks"], d.get("ivectors"))
PY
