#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c7
timeout -k 10 600 python -m pytest tests/test_gmm_gpu.py -x -q -m gpu -k "traintarget or full_size_oracle" 2>&1 | tail -n 6
timeout -k 10 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O.bench.json 2> $O.bench.err; echo "rc=$?"
python - <<'PY'
import json
j = json.loads(open('gpurun_out/r2c7.bench.json').read().strip().splitlines()[-1])
print('value %.1f M  e2e %.1f M  frac %.3f kernel_frac %.3f' % (j['value']/1e6, j['e2e']['value']/1e6, j['roofline']['frac'], j['roofline']['kernel_frac']))
for k in ('ivectors','strong_scaling','ivector_pipeline','tv_em'):
    print(k, json.dumps(j.get(k))[:600])
PY
tail -n 5 $O.bench.err
