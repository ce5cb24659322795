#!/bin/bash
# GPU call U: fused SEANet stage 1 -- parity, then the codec branch alone with and without it at 256 / 2,000 streams.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mimi.py tests/test_gpu_real_audio.py tests/test_gpu_asr.py tests/test_gpu_full_size.py -q -x -p no:cacheprovider > gpurun_out/u_tests.log 2>&1; tail -4 gpurun_out/u_tests.log
for f in 1 0; do
  for b in 256 2000; do
    echo "== fused1=$f streams=$b" >> gpurun_out/u_mimi.log
    DSM_SEANET_FUSED1=$f timeout 300 python tools/profile_steps.py --workload mimi --streams $b --steps 3 2>&1 | tail -1 >> gpurun_out/u_mimi.log
  done
done
cat gpurun_out/u_mimi.log
timeout 600 python bench.py > gpurun_out/u_bench256.json 2> gpurun_out/u_bench256.err; cat gpurun_out/u_bench256.json
DSM_SEANET_FUSED1=0 timeout 600 python bench.py > gpurun_out/u_bench256_off.json 2> gpurun_out/u_bench256_off.err; cat gpurun_out/u_bench256_off.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"seanet_stage1|init_conv|tc6" -c 60 --csv --log-file gpurun_out/u_launches.csv \
    python tools/profile_steps.py --workload mimi --eager --steps 2 > gpurun_out/u_ncu.log 2>&1
echo "ncu rc=$?"
