#!/bin/bash
# GPU call L: smoke(), launch list of the LM branch at 2,000 streams after the split fix, ncu --set full of the bf16
# attention, the four LM GEMM shapes and the codec kernels at 256 streams.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/l_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/l_smoke.log
timeout 300 python tools/profile_steps.py --workload lm --eager --steps 2 --streams 2000 --kv int8 > gpurun_out/l_plain_lm2000.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/l_launches_lm2000.csv \
    python tools/profile_steps.py --workload lm --eager --steps 2 --streams 2000 --kv int8 > gpurun_out/l_ncu1.log 2>&1
echo "ncu1 rc=$?"
timeout 300 python tools/profile_steps.py --workload lm --eager --steps 3 > gpurun_out/l_plain_lm256.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lm_attention_kernel" -s 16 -c 3 -o gpurun_out/r02l_attn_bf16 \
    python tools/profile_steps.py --workload lm --eager --steps 3 > gpurun_out/l_ncu2.log 2>&1
echo "ncu2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel" -s 64 -c 5 -o gpurun_out/r02l_lm_gemm \
    python tools/profile_steps.py --workload lm --eager --steps 3 > gpurun_out/l_ncu3.log 2>&1
echo "ncu3 rc=$?"
timeout 300 python tools/profile_steps.py --workload mimi --eager --steps 3 > gpurun_out/l_plain_mimi.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc6|codec_attention|rvq_select|init_conv" -s 160 -c 26 -o gpurun_out/r02l_codec \
    python tools/profile_steps.py --workload mimi --eager --steps 3 > gpurun_out/l_ncu4.log 2>&1
echo "ncu4 rc=$?"
ls -la gpurun_out/*.ncu-rep
