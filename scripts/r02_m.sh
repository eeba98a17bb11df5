set -x
python -m vina_slam_b200.replay --workload robosense128 --scans 40 --cold-start --out gpurun_out/r02_cold_start_robosense128_traj.txt > gpurun_out/r02_cold_start.log 2>&1; echo rc=$?; tail -3 gpurun_out/r02_cold_start.log
python bench.py --steps 20 --warmup 3 --no-cpu --batch 0 > gpurun_out/r02_bench_m.json 2> gpurun_out/r02_bench_m.err; echo rc=$?; tail -3 gpurun_out/r02_bench_m.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_m.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d['stage_ms'])"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "deskew or end_to_end or full_size" 2>&1 | tail -3
