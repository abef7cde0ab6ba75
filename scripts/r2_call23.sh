#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c23
timeout -k 10 900 python -m pytest tests/test_multi_gpu.py tests/test_multi_gpu_cli.py -x -q -m gpu > $O.pytest2.log 2>&1; echo "rc=$?" >> $O.pytest2.log
tail -n 5 $O.pytest2.log
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > $O.bench2.json 2> $O.bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2c23.bench2.json") if l.startswith("{")][-1])
for k in ("value","ms_per_step","e2e","strong_scaling","ivectors","ivector_pipeline","tv_em","plda"):
    print(k, json.dumps(d.get(k))[:420])
PY
tail -n 3 $O.bench2.err
