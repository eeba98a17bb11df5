set -x
python bench.py --steps 4 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02b_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 120 --csv --log-file gpurun_out/r02b_launches_final.csv python bench.py --steps 4 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02b_ncu_list.log 2>&1; echo ncu list rc=$?
ncu --set full --clock-control none --import-source on -k 'regex:k_iekf|k_insert_accum|k_recut_all|k_split|k_margi_leaves|k_deskew_var_init|k_insert_root|k_insert_leaf' -s 110 -c 11 -o gpurun_out/r02b_prof_final python bench.py --steps 4 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02b_ncu_full.log 2>&1; echo ncu full rc=$?
VINA_IEKF_LOOP=1 ncu --set full --clock-control none --import-source on -k 'regex:k_iekf_loop|k_down_emit_all|k_deskew_var_init_down' -s 12 -c 6 -o gpurun_out/r02b_prof_loop python bench.py --steps 4 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02b_ncu_full_loop.log 2>&1; echo ncu loop rc=$?
ncu --set full --clock-control none --import-source on -k 'regex:k_front_' -s 9 -c 3 -o gpurun_out/r02b_prof_front python scripts/front_end_time.py > gpurun_out/r02b_ncu_front.log 2>&1; echo ncu front rc=$?
ls -la gpurun_out/*.ncu-rep
