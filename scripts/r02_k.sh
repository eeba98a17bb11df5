set -x
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/r02_bench_k_n2.json 2> gpurun_out/r02_bench_k_n2.err; echo rc=$?; tail -12 gpurun_out/r02_bench_k_n2.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_k_n2.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step']); print(d.get('per_rank')); print(d.get('sharded')); print(d.get('clocks'))"
