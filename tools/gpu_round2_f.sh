#!/bin/bash
# 8-GPU call: the default bench line (cfg4 strong), sharded-filter tests at 3 / 4 / 8 ranks, cfg3 and cfg5 on 8 GPUs.
O=gpurun_out/r2; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 8 > $O/bench_cfg4_n8.json 2> $O/bench_cfg4_n8.err; echo "bench cfg4 n8 rc $?"
timeout 400 $TR --master-port 29522 bench.py --gpus 8 --workload cfg5 --steps 3 > $O/bench_cfg5_n8.json 2> $O/bench_cfg5_n8.err; echo "bench cfg5 n8 rc $?"
timeout 300 $TR --master-port 29523 bench.py --gpus 8 --workload cfg3 --steps 3 > $O/bench_cfg3_n8.json 2> $O/bench_cfg3_n8.err; echo "bench cfg3 n8 rc $?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 --no-latency > $O/bench_cfg4_n4.json 2> $O/bench_cfg4_n4.err; echo "bench cfg4 n4 rc $?"
timeout 600 python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q > $O/pytest_multigpu_n8.log 2>&1; echo "pytest rc $?"; tail -3 $O/pytest_multigpu_n8.log
python - <<PY
import json
for f in ("bench_cfg4_n8", "bench_cfg4_n4", "bench_cfg5_n8", "bench_cfg3_n8"):
    try:
        d = json.loads([l for l in open("$O/%s.json" % f) if l.startswith("{")][-1]); print(f, d["ms_per_step"], d["value"], d["e2e"]["value"], str(d.get("parity"))[:300])
    except Exception as e:
        print(f, "ERR", e)
PY
