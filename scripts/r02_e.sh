set -x
timeout 2400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_pytest_gpu_e.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/r02_pytest_gpu_e.log
python bench.py --steps 20 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err; echo rc=$?; tail -3 gpurun_out/r02_bench_e.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_e.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d['stage_ms'])"
python bench.py --steps 4 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_plain_e.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_launches_e.csv python bench.py --steps 4 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_ncu_e.log 2>&1; echo ncu rc=$?
ncu --set full --clock-control none --import-source on -k 'regex:k_iekf|k_insert_accum|k_recut_all|k_split|k_margi_leaves|k_recut_collect|k_margi_finish' -s 114 -c 11 -o gpurun_out/r02_prof_e python bench.py --steps 4 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_ncu_full_e.log 2>&1; echo ncu full rc=$?
ls -la gpurun_out/*.ncu-rep
