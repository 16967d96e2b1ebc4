rm -f gpurun_out/k2grid.txt
for e in 2 4 8 16 32; do DGB_PCG_K2_ELEMS=$e timeout 300 python tools/pcg_stage_times.py 128 256 512 2>&1 | grep "forward  auto" | sed "s/^/elems=$e /" >> gpurun_out/k2grid.txt; done
