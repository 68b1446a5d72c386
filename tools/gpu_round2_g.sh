#!/bin/bash
# usage: gpu_round2_g.sh <N>   -- A/B of the chunk-dealt global schedule at N GPUs
N=$1; O=gpurun_out/r2; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then timeout 300 python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q > $O/pytest_multigpu_deal.log 2>&1; echo "pytest rc $?"; tail -2 $O/pytest_multigpu_deal.log; fi
port=29530
for c in -1 0 512 16384; do
  port=$((port+1))
  timeout 300 $TR --master-port $port bench.py --gpus $N --no-latency --steps 5 --deal-chunk $c > $O/bench_cfg4_n${N}_deal$c.json 2> $O/bench_cfg4_n${N}_deal$c.err; echo "deal $c rc $?"
done
python - <<PY
import json
for c in (-1, 0, 512, 16384):
    try:
        d = json.loads([l for l in open("$O/bench_cfg4_n${N}_deal%d.json" % c) if l.startswith("{")][-1])
        p = d["parity"]
        print("chunk", c, "ms/step %.2f" % d["ms_per_step"], "kernel/rank", [round(v, 2) for v in d["per_rank"]["weighting_kernel_ms"]], "update/rank", [round(v, 2) for v in d["per_rank"]["update_ms"]],
              "parity", p.get("cloud_weight_max_rel_err"), p.get("normalised_w_bit_exact"), p.get("resample_indices_equal"))
    except Exception as e:
        print(c, "ERR", e)
PY
