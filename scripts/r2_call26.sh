#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c26
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > $O.bench8.json 2> $O.bench8.err; echo "bench8 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2c26.bench8.json") if l.startswith("{")][-1])
for k in ("value","ms_per_step","e2e","strong_scaling","ivectors","ivector_pipeline","tv_em","plda"):
    print(k, json.dumps(d.get(k))[:460])
PY
tail -n 3 $O.bench8.err
nvidia-smi topo -m > $O.topo.log 2>&1
