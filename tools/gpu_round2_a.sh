#!/bin/bash
# First GPU call of a session: GPU tests, the default bench line, the reference arm, the ncu launch list of the bench
# command and one `--set full` capture of the weighting kernel at cfg4 and cfg2.  Outputs under gpurun_out/r2/.
O=gpurun_out/r2; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench_cfg4_n1.json 2> $O/bench_cfg4_n1.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference > $O/bench_ref_n1.json 2> $O/bench_ref_n1.err; echo "ref rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg4.csv \
    python bench.py --steps 2 --warmup 3 --no-parity --no-latency > $O/launches_cfg4.out 2>&1; echo "ncu list rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:weight_v5 -s 20 -c 1 -f -o $O/weight_v5_cfg4 \
    python tools/prof_run.py cfg4 --updates 2 > $O/ncu_cfg4.out 2>&1; echo "ncu cfg4 rc $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:weight_v5 -s 2 -c 1 -f -o $O/weight_v5_cfg2 \
    python tools/prof_run.py cfg2 --updates 3 > $O/ncu_cfg2.out 2>&1; echo "ncu cfg2 rc $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:weight_v5 -s 2 -c 1 -f -o $O/weight_v5_cfg2_fast \
    python tools/prof_run.py cfg2 --updates 3 --opt reference_order=0 --opt sum_mode=2 > $O/ncu_cfg2_fast.out 2>&1; echo "ncu cfg2 fast rc $?"
ls -la $O
