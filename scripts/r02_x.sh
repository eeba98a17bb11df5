VINA_TRACE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --batch 0 2>&1 | grep -A4 "ktrace" | head -40
