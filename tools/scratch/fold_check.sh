set -x
timeout 900 python -m pytest tests/test_gpu_walker.py -x -q -k "pcg or slab" 2>&1 | tail -8 > gpurun_out/fold_tests.log
timeout 300 python -m pytest tests/test_gpu_dist.py tests/test_gpu_elliptic.py tests/test_gpu_multigrid.py -x -q 2>&1 | tail -5 >> gpurun_out/fold_tests.log
timeout 300 python tools/pcg_stage_times.py 512 1024 > gpurun_out/fold_stage_times.txt 2>&1
DGB_PCG_NO_FOLD=1 timeout 300 python tools/pcg_stage_times.py 1024 > gpurun_out/nofold_stage_times.txt 2>&1
timeout 300 python bench.py --no-micro --no-toefl --no-cpu-baseline > gpurun_out/fold_bench.json 2> gpurun_out/fold_bench.err
