timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --batch 0 2>&1 | grep "loop trace" | tail -3
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_s.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/r02_pytest_gpu_s.log
