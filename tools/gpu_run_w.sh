#!/bin/bash
# GPU call W: fused SEANet stage 1 (second version) parity + timing; launch list of the grouped RVQ; --set full of both.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_variants.py -k "FUSED1" -q -x -p no:cacheprovider > gpurun_out/w_tests.log 2>&1; tail -4 gpurun_out/w_tests.log
for f in 1 0; do
  for b in 256 2000; do
    echo "== fused1=$f streams=$b" >> gpurun_out/w_mimi.log
    DSM_SEANET_FUSED1=$f timeout 300 python tools/profile_steps.py --workload mimi --streams $b --steps 3 2>&1 | tail -1 >> gpurun_out/w_mimi.log
  done
done
cat gpurun_out/w_mimi.log
export DSM_SEANET_FUSED1=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rvq_chain|rvq_select|tc6|seanet_stage1|codec_attention" -c 150 --csv --log-file gpurun_out/w_launches.csv \
    python tools/profile_steps.py --workload mimi --eager --steps 2 > gpurun_out/w_ncu.log 2>&1
echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rvq_chain|seanet_stage1" -c 8 -o gpurun_out/r02w_stage1_rvq \
    python tools/profile_steps.py --workload mimi --eager --steps 1 > gpurun_out/w_ncu2.log 2>&1
echo "ncu2 rc=$?"
