VINA_TRACE=1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --batch 0 2>&1 | grep -A46 "ktrace\] k_split" | head -50 | cut -c1-250
