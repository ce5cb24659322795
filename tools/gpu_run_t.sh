#!/bin/bash
# GPU call T: new codec-ring test; re-capture of the q8 attention (the kernel file changed after call E).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mimi.py -q -s -p no:cacheprovider > gpurun_out/t_tests.log 2>&1; tail -4 gpurun_out/t_tests.log
timeout 300 python tools/profile_steps.py --workload lm --eager --steps 3 --kv int8 > gpurun_out/t_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lm_attention_q8 -s 16 -c 3 -o gpurun_out/r02t_attn_q8 \
    python tools/profile_steps.py --workload lm --eager --steps 3 --kv int8 > gpurun_out/t_ncu.log 2>&1
echo "ncu rc=$?"
timeout 300 python tools/profile_steps.py --workload mimi --eager --steps 3 > gpurun_out/t_plain_mimi.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc6|init_conv" -s 80 -c 14 -o gpurun_out/r02t_seanet \
    python tools/profile_steps.py --workload mimi --eager --steps 3 > gpurun_out/t_ncu2.log 2>&1
echo "ncu2 rc=$?"
