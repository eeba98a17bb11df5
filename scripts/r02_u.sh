ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_deskew_var_init|k_down_emit_all|k_iekf_loop" -c 18 --csv --log-file gpurun_out/r02_launches_u.csv python bench.py --steps 3 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_ncu_u.log 2>&1; echo rc=$?
python - <<'P'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/r02_launches_u.csv') if l.startswith('"')))
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size'); bi=h.index('Block Size')
seq=[(r[ki][:44],r[gi],r[bi],float(r[vi].replace(',',''))) for r in rows[1:]]
for s in seq[-30:]: print(s)
P
