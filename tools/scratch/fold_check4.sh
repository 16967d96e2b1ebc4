timeout 900 python -m pytest tests/test_gpu_walker.py -x -q 2>&1 | tail -3 > gpurun_out/fold_tests.log
rm -f gpurun_out/fold_w.txt
for w in 1.2 1.5 2.0 2.5 3.0 4.0; do DGB_WALK_SLOW_WEIGHT=$w timeout 300 python tools/pcg_stage_times.py 1024 2>&1 | grep "auto" | sed "s/^/w=$w /" >> gpurun_out/fold_w.txt; done
timeout 300 python tools/pcg_stage_times.py 512 2>&1 | grep "auto" >> gpurun_out/fold_w.txt
