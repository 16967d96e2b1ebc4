set -x
timeout 900 python -m pytest tests/test_gpu_walker.py -x -q -k "pcg or slab" 2>&1 | tail -3 > gpurun_out/fold_tests.log
timeout 300 python tools/pcg_stage_times.py 1024 > gpurun_out/fold_stage_times.txt 2>&1
for w in 1.5 2.5 3.0; do DGB_WALK_SLOW_WEIGHT=$w timeout 300 python tools/pcg_stage_times.py 1024 2>&1 | grep "auto" | sed "s/^/w=$w /" >> gpurun_out/fold_stage_times.txt; done
