#!/bin/bash
# compute-sanitizer memcheck over the address-arithmetic tests, cfg5 (HBM regime) bench line + ncu capture, cfg3 line.
O=gpurun_out/r2; mkdir -p $O
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9 \
    python -m pytest tests/test_gpu_weight.py tests/test_gpu_layout.py tests/test_gpu_filter.py tests/test_gpu_grid.py -m gpu -x -q \
    > $O/memcheck_1gpu.log 2>&1; echo "memcheck rc $?"; tail -4 $O/memcheck_1gpu.log
timeout 900 python bench.py --workload cfg5 --particles 1048576 --steps 5 --warmup 3 --no-latency > $O/bench_cfg5_1M_n1.json 2> $O/bench_cfg5_1M_n1.err; echo "cfg5 rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:weight_v5 -s 5 -c 1 -f -o $O/weight_v5_cfg5 \
    python tools/prof_run.py cfg5 --particles 1048576 --updates 1 > $O/ncu_cfg5.out 2>&1; echo "ncu cfg5 rc $?"
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 > $O/bench_cfg3_n1.json 2> $O/bench_cfg3_n1.err; echo "cfg3 rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:df_tile -c 1 -f -o $O/df_tile_mapL \
    python tools/prof_run.py cfg4 --grid-only > $O/ncu_df.out 2>&1; echo "ncu df rc $?"
ls -la $O | tail -12
