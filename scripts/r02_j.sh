set -x
( time python bench.py --steps 20 --warmup 3 ) > gpurun_out/r02_bench_j.json 2> gpurun_out/r02_bench_j.err; echo rc=$?; tail -6 gpurun_out/r02_bench_j.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_j.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d['stage_ms']); print(d.get('batch')); print(d.get('cpu_baseline'))"
