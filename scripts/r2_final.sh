#!/bin/bash
# final 1-GPU validation of the round: full GPU suite, smoke(), default bench (both arms), launch list of the bench
mkdir -p gpurun_out
O=gpurun_out/r2final
timeout -k 10 1500 python -m pytest tests -q -m gpu > $O.pytest.log 2>&1; echo "rc=$?" >> $O.pytest.log
tail -n 4 $O.pytest.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > $O.smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 $O.smoke.log
timeout -k 10 900 python bench.py > $O.bench.json 2> $O.bench.err; echo "bench rc=$?"
timeout -k 10 900 python bench.py --impl reference > $O.bench_ref.json 2> $O.bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2final.bench.json") if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "kernel", d["roofline"]["kernel_frac"], "launches", d["gpu_launches"])
print("e2e", d["e2e"]["value"], "clocks", d["clocks"])
for k in ("ivectors","ivector_pipeline","tv_em","plda","cpu_baseline"):
    print(k, json.dumps(d.get(k))[:300])
r=json.loads([l for l in open("gpurun_out/r2final.bench_ref.json") if l.startswith("{")][-1])
print("ref", r.get("value"), r.get("cpu_baseline"))
PY
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O.launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O.ncu_bench.log 2>&1; echo "ncu rc=$?"
