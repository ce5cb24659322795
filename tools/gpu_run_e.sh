#!/bin/bash
# GPU call E: fused RVQ after the winner-row fix, session test, live sweep with the int8 ring, ncu capture of the q8 attention.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mimi.py tests/test_gpu_real_audio.py tests/test_gpu_session.py tests/test_gpu_sched.py tests/test_gpu_asr.py -q -s -p no:cacheprovider > gpurun_out/e_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/e_tests.log
tail -6 gpurun_out/e_tests.log
for mx in 512 0; do
  DSM_RVQ_FUSED_MAX=$mx timeout 300 python tools/profile_steps.py --workload mimi --steps 5 > gpurun_out/e_mimi_fused$mx.txt 2>&1
  grep -h alone gpurun_out/e_mimi_fused$mx.txt
  DSM_RVQ_FUSED_MAX=$mx timeout 300 python tools/profile_steps.py --workload mimi --steps 5 --streams 64 > gpurun_out/e_mimi64_fused$mx.txt 2>&1
  grep -h alone gpurun_out/e_mimi64_fused$mx.txt
done
timeout 900 python bench.py --steps 20 --warmup 5 --kv int8 --streams 2000 --latency-steps 200 --no-cpu-baseline --live-sweep 2400,2800 \
   > gpurun_out/e_bench_2000_int8_sweep.json 2> gpurun_out/e_bench_2000_int8_sweep.err
echo "sweep rc=$?"
timeout 300 python tools/profile_steps.py --eager --steps 3 --kv int8 > gpurun_out/e_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lm_attention_q8 -s 16 -c 3 -o gpurun_out/r02e_attn_q8 \
    python tools/profile_steps.py --eager --steps 3 --kv int8 > gpurun_out/e_ncu.log 2>&1
echo "ncu rc=$?"
