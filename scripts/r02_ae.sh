for i in 1 2 3 4; do timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_bench_ae$i.json 2> gpurun_out/r02_bench_ae$i.err; echo rc=$?; tail -2 gpurun_out/r02_bench_ae$i.err; python -c "
import json
d=json.load(open('gpurun_out/r02_bench_ae$i.json')); print(round(d['ms_per_step'],4), round(d['ms_per_step_median'],4), round(d['ms_per_step_max'],4), '|', round(d['e2e']['ms_per_step'],4), round(d['e2e']['ms_per_step_median'],4), round(d['e2e']['ms_per_step_max'],4), d['gpu_launches'], d['clocks'])
"; done
