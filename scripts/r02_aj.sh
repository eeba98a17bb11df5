timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_aj.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r02_pytest_gpu_aj.log
for E in 1 0 1 0; do VINA_EARLY_COLLECT=$E timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_bench_aj$E.json 2> gpurun_out/r02_bench_aj$E.err; python -c "
import json
d=json.load(open('gpurun_out/r02_bench_aj$E.json')); print($E, round(d['ms_per_step'],4), round(d['ms_per_step_median'],4), '|', round(d['e2e']['ms_per_step'],4), round(d['e2e']['ms_per_step_median'],4), d['gpu_launches'])
"; done
