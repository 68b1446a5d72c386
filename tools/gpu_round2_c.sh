#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_weight_stress.py tests/test_gpu_layout.py tests/test_gpu_parity_benched.py -m gpu -x -q > $O/pytest_ordered.log 2>&1; echo "pytest rc $?"; tail -15 $O/pytest_ordered.log
timeout 600 python tools/sweep.py cfg2 --spec tools/specs/r4a_ordered.json --out $O/sweep_r4a_cfg2.jsonl > $O/sweep_r4a_cfg2.log 2>&1; echo "sweep cfg2 rc $?"
timeout 300 python tools/sweep.py cfg1 --spec tools/specs/r4a_ordered.json --out $O/sweep_r4a_cfg1.jsonl > $O/sweep_r4a_cfg1.log 2>&1; echo "sweep cfg1 rc $?"
timeout 600 python tools/sweep.py cfg5 --spec tools/specs/r4a_ordered.json --steps 2 --out $O/sweep_r4a_cfg5.jsonl > $O/sweep_r4a_cfg5.log 2>&1; echo "sweep cfg5 rc $?"
cat $O/sweep_r4a_cfg2.jsonl $O/sweep_r4a_cfg1.jsonl $O/sweep_r4a_cfg5.jsonl | cut -c1-400
