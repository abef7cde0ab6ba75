#!/bin/bash
for k in 3 2; do
timeout -k 5 200 python bench.py --kernel $k --frames 10000000 --steps 3 --warmup 3 --no-cpu-baseline --no-ivectors --e2e-steps 3 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read())
print('kernel', j['config']['kernel'], 'value %.1f M' % (j['value']/1e6), 'e2e %.1f M' % (j['e2e']['value']/1e6), 'ms', j['ms_per_step'])
"
done
nvidia-smi topo -m 2>/dev/null | head -5
