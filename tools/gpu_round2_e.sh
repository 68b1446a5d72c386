#!/bin/bash
# 2-GPU call: sharded-filter tests against the oracle, the default bench line at N = 2 (cfg4 strong) and cfg2 strong
# (exercises the fused ordered kernel under the pose-balanced global schedule), memcheck of the 2-rank peer-mailbox path.
O=gpurun_out/r2; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q > $O/pytest_multigpu_n2.log 2>&1; echo "pytest rc $?"; tail -3 $O/pytest_multigpu_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $O/bench_cfg4_n2.json 2> $O/bench_cfg4_n2.err; echo "bench n2 rc $?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload cfg2 > $O/bench_cfg2_n2.json 2> $O/bench_cfg2_n2.err; echo "bench cfg2 n2 rc $?"
timeout 400 compute-sanitizer --target-processes all --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q -k "2-uneven or 2-empty" > $O/memcheck_2gpu.log 2>&1; echo "memcheck2 rc $?"; tail -5 $O/memcheck_2gpu.log
python - <<PY
import json
for f in ("$O/bench_cfg4_n2.json", "$O/bench_cfg2_n2.json"):
    try:
        d = json.load(open(f)); print(f, d["ms_per_step"], d["value"], d["e2e"]["value"], d["parity"])
    except Exception as e:
        print(f, "ERR", e)
PY
