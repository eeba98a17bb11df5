# ncu captures of the kernels added at the end of round 1 (scan front end, map pruning) - to be run under gpurun at
# the start of round 2, each only after the plain command has exited 0 (B200_PROFILING.md):
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash scripts/profile_front_and_prune.sh'
set -x
python scripts/front_end_time.py > gpurun_out/front_end_time.json 2> gpurun_out/front_end_time.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_front.csv \
    python scripts/front_end_time.py > gpurun_out/ncu_front.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sort_scatter -c 4 -o gpurun_out/k_sort_scatter \
    python scripts/front_end_time.py > gpurun_out/ncu_front_full.log 2>&1
python bench.py --mode bigmap --workload hilti_xt32 --voxels 1e6 --steps 10 --warmup 3 > gpurun_out/bigmap_prune_1e6.json \
    2> gpurun_out/bigmap_prune_1e6.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_prune_sweep -c 2 -o gpurun_out/k_prune_sweep \
    python bench.py --mode bigmap --workload hilti_xt32 --voxels 1e6 --steps 10 --warmup 3 > gpurun_out/ncu_prune_full.log 2>&1
