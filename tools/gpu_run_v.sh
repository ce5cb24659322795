#!/bin/bash
# GPU call V: grouped (Gram-table) RVQ encode -- parity, then the codec branch alone per group size.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_mimi.py tests/test_gpu_real_audio.py tests/test_gpu_asr.py tests/test_gpu_full_size.py tests/test_gpu_variants.py -q -x -p no:cacheprovider > gpurun_out/v_tests.log 2>&1; tail -6 gpurun_out/v_tests.log
for g in 8 4 2 1; do
  for b in 64 256 2000; do
    echo "== rvq_group=$g streams=$b" >> gpurun_out/v_mimi.log
    DSM_RVQ_GROUP=$g timeout 300 python tools/profile_steps.py --workload mimi --streams $b --steps 3 2>&1 | tail -1 >> gpurun_out/v_mimi.log
  done
done
cat gpurun_out/v_mimi.log
timeout 600 python bench.py > gpurun_out/v_bench256.json 2> gpurun_out/v_bench256.err; cat gpurun_out/v_bench256.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rvq|tc6" -c 120 --csv --log-file gpurun_out/v_launches.csv \
    python tools/profile_steps.py --workload mimi --eager --steps 2 > gpurun_out/v_ncu.log 2>&1
echo "ncu rc=$?"
