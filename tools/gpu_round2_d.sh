#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
for w in cfg2 cfg1; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_$w.csv python tools/prof_run.py $w --updates 4 > $O/launches_$w.out 2>&1; echo "$w rc $?"
python tools/launch_summary.py $O/launches_$w.csv
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:weight_ordered -s 1 -c 1 -f -o $O/weight_ordered_cfg2 python tools/prof_run.py cfg2 --updates 3 > $O/ncu_ordered.out 2>&1; echo "ncu ordered rc $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:update_seg -s 1 -c 1 -f -o $O/update_seg_cfg2 python tools/prof_run.py cfg2 --updates 3 > $O/ncu_useg.out 2>&1; echo "ncu useg rc $?"
