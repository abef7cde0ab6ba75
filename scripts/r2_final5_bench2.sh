#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2final5_bench2
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > $O.json 2> $O.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2final5_bench2.json") if l.startswith("{")][-1])
print("value", d["value"], "n", d["n_gpus"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
for k in ("strong_scaling","ivectors","ivector_pipeline","tv_em","plda","product_levels"):
    print(k, json.dumps(d.get(k))[:200])
PY
