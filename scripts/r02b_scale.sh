set -x
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 12 --warmup 3 > gpurun_out/r02b_bench_n$N.json 2> gpurun_out/r02b_bench_n$N.err; echo N$N rc=$?; tail -3 gpurun_out/r02b_bench_n$N.err
python -c "
import json
d=json.load(open('gpurun_out/r02b_bench_n$N.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d.get('per_rank',{}).get('ms_per_step')); s=d.get('sharded'); print(json.dumps(s)[:1500] if s else None)
"
