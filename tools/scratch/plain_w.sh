rm -f gpurun_out/plain_w.txt
for w in 1.8 2.3 3.0 4.0 5.0; do DGB_WALK_SLOW_WEIGHT=$w timeout 300 python tools/microbench.py 1024 15 2>&1 | grep -E "Elliptic2d (forward|centered) FUSED" | sed "s/^/w=$w /" >> gpurun_out/plain_w.txt; done
