set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo N8 rc=$?; tail -3 gpurun_out/r02_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --mode sharded-map --p2p --query --workload hilti_xt32 --voxels 1e7 --steps 10 --warmup 3 > gpurun_out/r02_sharded_hilti_1e7_per_rank_n8.json 2> gpurun_out/r02_sharded_hilti_1e7_per_rank_n8.err; echo shard rc=$?; tail -3 gpurun_out/r02_sharded_hilti_1e7_per_rank_n8.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_n8.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d.get('per_rank')); print(d.get('sharded'))
d=json.load(open('gpurun_out/r02_sharded_hilti_1e7_per_rank_n8.json')); print(json.dumps(d)[:1800])
"
