set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_r.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/r02_pytest_gpu_r.log
for L in 1 0; do VINA_IEKF_LOOP=$L VINA_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_bench_loop$L.json 2> gpurun_out/r02_bench_loop$L.err; echo loop$L rc=$?; tail -4 gpurun_out/r02_bench_loop$L.err; done
python -c "
import json
for f in ['r02_bench_loop1','r02_bench_loop0']:
    d=json.load(open('gpurun_out/'+f+'.json')); print(f, round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['roofline'], d['stage_ms'], d['config'].get('gt_traj_err_m'))
"
