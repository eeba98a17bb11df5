set -x
timeout 2400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_pytest_gpu_i.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/r02_pytest_gpu_g.log
python bench.py --steps 20 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_bench_i.json 2> gpurun_out/r02_bench_i.err; echo rc=$?; tail -3 gpurun_out/r02_bench_i.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_i.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d['stage_ms'])"
python bench.py --steps 4 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_plain_i.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_launches_i.csv python bench.py --steps 4 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_ncu_i.log 2>&1; echo ncu rc=$?
