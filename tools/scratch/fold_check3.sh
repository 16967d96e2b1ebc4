for w in 1.8 2.2 3.5 4.0 5.0; do DGB_WALK_SLOW_WEIGHT=$w timeout 300 python tools/pcg_stage_times.py 1024 2>&1 | grep "auto" | sed "s/^/w=$w /" >> gpurun_out/fold_w.txt; done
for o in 0.5 4.0; do DGB_WALK_TASK_OVERHEAD=$o timeout 300 python tools/pcg_stage_times.py 1024 2>&1 | grep "auto" | sed "s/^/ovh=$o /" >> gpurun_out/fold_w.txt; done
for w in 2.0 3.0 4.0; do DGB_WALK_SLOW_WEIGHT=$w timeout 300 python tools/pcg_stage_times.py 512 2>&1 | grep "auto" | sed "s/^/w=$w /" >> gpurun_out/fold_w.txt; done
