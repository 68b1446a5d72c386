#!/bin/bash
N=$1; shift; O=gpurun_out/r2; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29560
for c in "$@"; do
  port=$((port+1))
  timeout 200 $TR --master-port $port bench.py --gpus $N --no-latency --no-parity --steps 5 --deal-chunk $c > $O/bench_cfg4_n${N}_p_deal$c.json 2> $O/bench_cfg4_n${N}_p_deal$c.err; echo "deal $c rc $?"
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/bench_cfg4_n${N}_p_deal$c.json") if l.startswith("{")][-1])
    r = lambda k: [round(v, 2) for v in d["per_rank"][k]]
    print("chunk $c ms/step %.2f" % d["ms_per_step"], "kernel", r("weighting_kernel_ms"), "exch+wait", r("exchange_and_wait_ms"), "sums", r("sums_over_particles_ms"), "fast %.2f" % d["fast_mode"]["ms_per_step"])
except Exception as e:
    print("$c", "ERR", e)
PY
done
