set -x
python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu_now.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_now.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_now.log
python bench.py --impl reference > gpurun_out/bench_ref_now.json 2> gpurun_out/bench_ref_now.err
python bench.py > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err
python tools/shim_toefl_bench.py > gpurun_out/shim_toefl_now.json 2> gpurun_out/shim_toefl_now.err
