for i in 1 2; do timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_bench_z$i.json 2> gpurun_out/r02_bench_z$i.err; echo rc=$?; tail -2 gpurun_out/r02_bench_z$i.err; python -c "
import json
d=json.load(open('gpurun_out/r02_bench_z$i.json')); print(round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['gpu_launches'], d['roofline']['launch_us'], d['roofline']['frac'], d['stage_ms']['iekf_ms'], d['config'].get('gt_traj_err_m'))
"; done
