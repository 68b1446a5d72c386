#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_filter.py tests/test_gpu_host_classes.py -m gpu -x -q > $O/pytest_filter.log 2>&1; echo "pytest rc $?"; tail -12 $O/pytest_filter.log
timeout 600 python bench.py --workload cfg5 --particles 1048576 --steps 3 --warmup 3 --no-latency > $O/bench_cfg5_1M_n1_b.json 2> $O/bench_cfg5_1M_n1_b.err; echo "cfg5 rc $?"
python - <<PY
import json
d = json.loads([l for l in open("$O/bench_cfg5_1M_n1_b.json") if l.startswith("{")][-1])
print(d["ms_per_step"], d["parity"])
PY
