timeout 900 python -m pytest tests -m gpu -x -q -k "map or overlapped or full_size or trajectory or insert" > gpurun_out/r02_pytest_gpu_al.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/r02_pytest_gpu_al.log
for i in 1 2 3; do timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_bench_al$i.json 2> gpurun_out/r02_bench_al$i.err; python -c "
import json
d=json.load(open('gpurun_out/r02_bench_al$i.json')); print(round(d['ms_per_step'],4), round(d['ms_per_step_median'],4), '|', round(d['e2e']['ms_per_step'],4), d['stage_ms']['insert_ms'], d['stage_ms']['recut_ms'])
"; done
