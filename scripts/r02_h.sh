set -x
VINA_TRACE=1 python bench.py --steps 6 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_bench_h.json 2> gpurun_out/r02_bench_h.err; echo rc=$?; grep -A4 "ktrace\|vina trace" gpurun_out/r02_bench_h.err | tail -60
