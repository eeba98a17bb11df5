for w in mid360 velodyne32; do python bench.py --workload $w --ba --steps 20 --warmup 9 --batch 0 > gpurun_out/r02b_bench_ba_$w.json 2> gpurun_out/r02b_bench_ba_$w.err; echo ba $w rc=$?; done
ncu --set full --clock-control none --import-source on -k 'regex:k_front_' -s 9 -c 3 -o gpurun_out/r02b_prof_front python scripts/front_end_time.py > gpurun_out/r02b_ncu_front.log 2>&1; echo ncu front rc=$?
