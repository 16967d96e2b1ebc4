set -x
python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_reference.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/launches_r02_bench_pcg12.csv python bench.py --steps 2 --warmup 3 --iters 12 --no-micro --no-toefl --no-cpu-baseline > gpurun_out/launches_r02.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'walker|pcg_update' -s 6 -c 2 -o gpurun_out/prof_r02_pcg_fold python tools/prof_kernels.py pcg > gpurun_out/prof_r02_pcg_fold.log 2>&1
python tools/microbench.py 1024 30 > gpurun_out/microbench_r02_1024.txt 2>&1
python tools/microbench.py 128 30 > gpurun_out/microbench_r02_config1_128.txt 2>&1
python tools/pcg_stage_times.py 128 256 512 1024 > gpurun_out/pcg_stage_times_r02.txt 2>&1
python bench.py --workload toefl --steps 6 > gpurun_out/toefl_r02.json 2>&1
python bench.py --workload ds > gpurun_out/ds_r02.json 2> gpurun_out/ds_r02.err
python tools/shim_toefl_bench.py > gpurun_out/shim_toefl_r02.json 2>&1
python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/pytest_gpu_r02.log
