set -x
timeout 1200 python -m pytest tests/test_gpu_init.py -m gpu -q -x > gpurun_out/r02_pytest_gpu_init.log 2>&1; echo pytest rc=$?; tail -40 gpurun_out/r02_pytest_gpu_init.log
