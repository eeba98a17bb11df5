for P in 1 0 1 0; do VINA_PDL=$P timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_bench_pdl$P.json 2> gpurun_out/r02_bench_pdl$P.err; echo pdl$P rc=$?; tail -2 gpurun_out/r02_bench_pdl$P.err; python -c "
import json
d=json.load(open('gpurun_out/r02_bench_pdl$P.json')); print($P, round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['gpu_launches'], d['roofline']['launch_us'], d['stage_ms'])
"; done
