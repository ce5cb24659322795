#!/bin/bash
# GPU call N: where the wide LM GEMM tile starts to pay (threshold on the number of 128 x 256 tiles).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for t in 1 100 148 300 1000000; do
  for b in 2000 1024 512; do
    DSM_TC_WIDE_MIN_TILES=$t timeout 300 python tools/profile_steps.py --workload lm --steps 5 --streams $b --kv int8 > gpurun_out/n_lm_${t}_$b.txt 2>&1
    echo "wide_min=$t streams=$b $(grep -h alone gpurun_out/n_lm_${t}_$b.txt)"
  done
done
