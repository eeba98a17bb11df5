set -x
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_d.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/r02_pytest_gpu_d.log
python bench.py --steps 20 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; echo rc=$?; tail -3 gpurun_out/r02_bench_d.err; head -c 2500 gpurun_out/r02_bench_d.json
python bench.py --steps 4 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_plain_d.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_launches_d.csv python bench.py --steps 4 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_ncu_d.log 2>&1; echo ncu rc=$?
