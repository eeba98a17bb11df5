set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest_gpu_final.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r02b_pytest_gpu_final.log
python bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/r02b_bench_reference.json 2> gpurun_out/r02b_bench_reference.err; echo ref rc=$?
python bench.py --steps 20 --warmup 3 > gpurun_out/r02b_bench_final.json 2> gpurun_out/r02b_bench_final.err; echo ours rc=$?
VINA_IEKF_LOOP=1 python bench.py --steps 20 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02b_bench_loop_schedule.json 2> gpurun_out/r02b_bench_loop_schedule.err; echo loop rc=$?
for w in mid360 velodyne32 hilti_xt32; do python bench.py --workload $w --steps 20 --warmup 3 --batch 0 > gpurun_out/r02b_bench_$w.json 2> gpurun_out/r02b_bench_$w.err; echo $w rc=$?; done
for w in mid360 velodyne32; do python bench.py --workload $w --ba --steps 20 --warmup 9 > gpurun_out/r02b_bench_ba_$w.json 2> gpurun_out/r02b_bench_ba_$w.err; echo ba $w rc=$?; done
python scripts/front_end_time.py > gpurun_out/r02b_front_end_time.json 2> gpurun_out/r02b_front_end_time.err; echo front rc=$?
python scripts/h2d_probe.py > gpurun_out/r02b_h2d_probe.txt 2>&1
python -c "
import json
for f in ['r02b_bench_final','r02b_bench_reference','r02b_bench_loop_schedule','r02b_bench_mid360','r02b_bench_velodyne32','r02b_bench_hilti_xt32','r02b_bench_ba_mid360','r02b_bench_ba_velodyne32']:
    try:
        d=json.load(open('gpurun_out/'+f+'.json')); print(f, round(d['ms_per_step'],4), round(d.get('e2e',{}).get('ms_per_step',0),4), d.get('cpu_baseline',{}).get('ms_per_step'), d.get('roofline',{}).get('frac'), d.get('batch',{}).get('value'), d.get('gpu_launches'))
    except Exception as e: print(f, 'ERR', e)
"
