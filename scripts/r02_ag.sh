timeout 600 python -m pytest tests -m gpu -x -q -k "loop or overlapped or retry or end_to_end" 2>&1 | tail -2
for L in 1; do VINA_TRACE=1 VINA_IEKF_LOOP=$L timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_bench_ag$L.json 2> gpurun_out/r02_bench_ag$L.err; echo loop$L rc=$?; grep "vina trace" gpurun_out/r02_bench_ag$L.err | sed 's/.*device us/device us/'; done
for L in 1 0 1 0; do VINA_IEKF_LOOP=$L timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_bench_ag$L.json 2> gpurun_out/r02_bench_ag$L.err; python -c "
import json
d=json.load(open('gpurun_out/r02_bench_ag$L.json')); print($L, round(d['ms_per_step'],4), round(d['ms_per_step_median'],4), '|', round(d['e2e']['ms_per_step'],4), d['gpu_launches'], d['stage_ms']['iekf_ms'])
"; done
