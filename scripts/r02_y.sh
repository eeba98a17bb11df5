for P in 1 0 1 0; do echo so$P; VINA_TRACE=1 VINA_SPLIT_OVERLAP=$P timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu --batch 0 2>&1 | grep "vina trace" | sed 's/.*device us/device us/'; done
