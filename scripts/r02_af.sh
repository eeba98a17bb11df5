VINA_IEKF_LOOP=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_bench_af.json 2> gpurun_out/r02_bench_af.err; python -c "
import json
d=json.load(open('gpurun_out/r02_bench_af.json')); print(round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['gpu_launches'], d['roofline']['launch_us'], d['stage_ms'])
"
VINA_IEKF_LOOP=1 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread --clock-control none -k regex:"k_iekf_loop" -c 6 --csv --log-file gpurun_out/r02_launches_af.csv python bench.py --steps 3 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_ncu_af.log 2>&1
grep k_iekf_loop gpurun_out/r02_launches_af.csv | awk -F'","' '{print $13, $15}'
