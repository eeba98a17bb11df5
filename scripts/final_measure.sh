set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r01_final.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu_r01_final.log
python bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/bench_r01b_reference.json 2> gpurun_out/bench_r01b_reference.err; echo ref rc=$?
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01b_final.json 2> gpurun_out/bench_r01b_final.err; echo ours rc=$?
for w in mid360 velodyne32 hilti_xt32; do python bench.py --workload $w --steps 20 --warmup 3 --batch 0 > gpurun_out/bench_r01b_$w.json 2> gpurun_out/bench_r01b_$w.err; echo $w rc=$?; done
for w in mid360 velodyne32; do python bench.py --workload $w --ba --steps 20 --warmup 9 > gpurun_out/bench_r01b_ba_$w.json 2> gpurun_out/bench_r01b_ba_$w.err; echo ba $w rc=$?; done
python bench.py --steps 4 --warmup 3 --no-cpu --batch 0 > gpurun_out/plain_r01b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 4 --warmup 3 --no-cpu --batch 0 > gpurun_out/ncu_r01b.log 2>&1; echo ncu rc=$?
python -c "
import json
for f in ['bench_r01b_final','bench_r01b_reference','bench_r01b_mid360','bench_r01b_velodyne32','bench_r01b_hilti_xt32','bench_r01b_ba_mid360','bench_r01b_ba_velodyne32']:
    try:
        d=json.load(open('gpurun_out/'+f+'.json')); print(f, round(d['ms_per_step'],4), round(d.get('e2e',{}).get('ms_per_step',0),4), d.get('cpu_baseline',{}).get('ms_per_step'), d.get('roofline',{}).get('frac'), d.get('batch',{}).get('value'))
    except Exception as e: print(f, 'ERR', e)
"
