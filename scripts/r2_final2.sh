#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2final2
timeout -k 10 1500 python -m pytest tests -q -m gpu > $O.pytest.log 2>&1; echo "rc=$?" >> $O.pytest.log
tail -n 4 $O.pytest.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
