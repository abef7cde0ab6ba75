#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c29
timeout -k 10 600 python scripts/products_probe.py > $O.probe.log 2>&1; echo "probe rc=$?"; cat $O.probe.log | cut -c1-600
for L in 0 1 2; do
  timeout -k 10 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra --no-ivectors --e2e-steps 1 --products $L > $O.bench$L.json 2> $O.bench$L.err; echo "bench$L rc=$?"
  python - $L <<'PY'
import json,sys
L=sys.argv[1]
d=json.loads([l for l in open(f"gpurun_out/r2c29.bench{L}.json") if l.startswith("{")][-1])
print("level",L,"value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "kernel", d["roofline"]["kernel_frac"], "llk", d["config"]["mean_llk_per_frame"], d["clocks"]["sm_mhz"])
PY
done
