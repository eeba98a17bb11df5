set -x
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 12 --warmup 3 > gpurun_out/r02_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err; echo N=$N rc=$?; tail -2 gpurun_out/r02_scale_n$N.err; head -c 600 gpurun_out/r02_scale_n$N.json; echo
done
