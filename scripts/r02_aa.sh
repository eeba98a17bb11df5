timeout 600 python -m pytest tests -m gpu -x -q -k "front or unsorted or raw" 2>&1 | tail -3
python scripts/front_end_time.py > gpurun_out/r02_front_end_time.json 2> gpurun_out/r02_front_end_time.err; echo rc=$?; cat gpurun_out/r02_front_end_time.json; tail -3 gpurun_out/r02_front_end_time.err
VINA_FRONT_RADIX=1 python scripts/front_end_time.py 2>&1 | tail -1
