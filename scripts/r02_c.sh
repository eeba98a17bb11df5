set -x
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k "large_map" > gpurun_out/r02_pytest_gpu_c.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/r02_pytest_gpu_c.log
python bench.py --steps 4 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_plain_c.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_launches_base.csv python bench.py --steps 4 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_ncu_c.log 2>&1; echo ncu rc=$?
