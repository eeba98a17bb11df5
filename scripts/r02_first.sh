set -x
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_a.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/r02_pytest_gpu_a.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo ours rc=$?; tail -3 gpurun_out/r02_bench_a.err
cat gpurun_out/r02_bench_a.json | head -c 3000
