#!/bin/bash
# Final single-GPU validation of a round: all GPU tests, smoke(), the default bench line.
O=gpurun_out/r2; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/final_pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 $O/final_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc $?"; tail -2 $O/final_smoke.log
timeout 600 python bench.py > $O/final_bench_cfg4_n1.json 2> $O/final_bench_cfg4_n1.err; echo "bench rc $?"
python - <<PY
import json
d = json.loads([l for l in open("$O/final_bench_cfg4_n1.json") if l.startswith("{")][-1])
print("ms/step %.2f value %.4g e2e %.4g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), d["per_rank"], d["parity"]["cloud_weight_max_rel_err"], d["parity"]["normalised_w_bit_exact"], d["parity"]["resample_indices_equal"], d["parity"]["mean_bit_exact_where_claimed"])
print({k: {m: v2.get("update_p50_ms") for m, v2 in v.items() if isinstance(v2, dict)} for k, v in d["latency"].items()})
PY
