set -x
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_l.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/r02_pytest_gpu_l.log
