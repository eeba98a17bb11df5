for combo in "768 1" "512 2" "384 2" "256 3" "256 4"; do
  set -- $combo
  export VINA_IEKF_THREADS=$1 VINA_IEKF_BLOCKS_PER_SM=$2
  touch vina_slam_b200/csrc/iekf_kernel.cu
  python -m vina_slam_b200.build -v 2>&1 | grep -A2 "Function properties for _Z6k_iekfILb0" | grep -E "Used" 
  python scripts/iekf_microbench.py > gpurun_out/iekf_micro_v8_$1x$2.txt 2>&1
  echo "== $combo"; grep -E "product|18 blocks  |37 blocks" gpurun_out/iekf_micro_v8_$1x$2.txt
done
