timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_bench_v.json 2> gpurun_out/r02_bench_v.err; echo rc=$?; tail -2 gpurun_out/r02_bench_v.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_v.json')); print(round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['gpu_launches'], d['roofline']['launch_us'], d['stage_ms'], d['config'].get('gt_traj_err_m'))
"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_deskew_var_init|k_down_emit_all|k_iekf_loop" -c 9 --csv --log-file gpurun_out/r02_launches_v.csv python bench.py --steps 3 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_ncu_v.log 2>&1; echo rc=$?
python - <<'P'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/r02_launches_v.csv') if l.startswith('"')))
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[1:]: print(r[ki][:40], r[vi])
P
