for i in 1 2; do VINA_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_bench_ac$i.json 2> gpurun_out/r02_bench_ac$i.err; echo rc=$?; grep "vina trace" gpurun_out/r02_bench_ac$i.err | sed 's/.*steps; //'; python -c "
import json
d=json.load(open('gpurun_out/r02_bench_ac$i.json')); print(round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['gpu_launches'])
"; done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_bench_ac3.json 2> gpurun_out/r02_bench_ac3.err; python -c "
import json
d=json.load(open('gpurun_out/r02_bench_ac3.json')); print(round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['gpu_launches'])
"
