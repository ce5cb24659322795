#!/bin/bash
# GPU call X: fused SEANet stage 1, third version (warp-uniform weights) -- parity, timing, one --set full capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_variants.py -k "FUSED1" -q -x -p no:cacheprovider > gpurun_out/x_tests.log 2>&1; tail -4 gpurun_out/x_tests.log
for f in 1 0; do
  for b in 64 256 2000; do
    echo "== fused1=$f streams=$b" >> gpurun_out/x_mimi.log
    DSM_SEANET_FUSED1=$f timeout 300 python tools/profile_steps.py --workload mimi --streams $b --steps 3 2>&1 | tail -1 >> gpurun_out/x_mimi.log
  done
done
cat gpurun_out/x_mimi.log
export DSM_SEANET_FUSED1=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"seanet_stage1" -c 2 -o gpurun_out/r02x_stage1 \
    python tools/profile_steps.py --workload mimi --eager --steps 1 > gpurun_out/x_ncu2.log 2>&1
echo "ncu2 rc=$?"
