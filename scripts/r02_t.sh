set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_t.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/r02_pytest_gpu_t.log
for C in "1 1" "1 0" "0 0"; do set -- $C; VINA_IEKF_LOOP=$1 VINA_FRONT_FUSED=$2 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_bench_l$1f$2.json 2> gpurun_out/r02_bench_l$1f$2.err; echo l$1f$2 rc=$?; tail -2 gpurun_out/r02_bench_l$1f$2.err; done
VINA_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --batch 0 2>&1 | grep "vina trace"
python -c "
import json
for f in ['r02_bench_l1f1','r02_bench_l1f0','r02_bench_l0f0']:
    d=json.load(open('gpurun_out/'+f+'.json')); print(f, round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['gpu_launches'], d['roofline']['launch_us'], d['stage_ms'], d['config'].get('gt_traj_err_m'))
"
