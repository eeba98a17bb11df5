set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest_gpu_final.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/r02b_pytest_gpu_final.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02b_bench_final.json 2> gpurun_out/r02b_bench_final.err; echo ours rc=$?
VINA_IEKF_LOOP=1 python bench.py --steps 20 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02b_bench_loop_schedule.json 2> gpurun_out/r02b_bench_loop_schedule.err; echo loop rc=$?
for w in mid360 velodyne32 hilti_xt32; do python bench.py --workload $w --steps 20 --warmup 3 --batch 0 > gpurun_out/r02b_bench_$w.json 2> gpurun_out/r02b_bench_$w.err; echo $w rc=$?; done
python bench.py --steps 4 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02b_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 60 --csv --log-file gpurun_out/r02b_launches_final.csv python bench.py --steps 4 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02b_ncu_list.log 2>&1; echo ncu list rc=$?
python -c "
import json
for f in ['r02b_bench_final','r02b_bench_loop_schedule','r02b_bench_mid360','r02b_bench_velodyne32','r02b_bench_hilti_xt32']:
    d=json.load(open('gpurun_out/'+f+'.json')); print(f, round(d['ms_per_step'],4), round(d['ms_per_step_median'],4), round(d['e2e']['ms_per_step'],4), d.get('cpu_baseline',{}).get('ms_per_step'), d['roofline']['frac'], d.get('batch',{}).get('value'), d['gpu_launches'], d['clocks'])
"
