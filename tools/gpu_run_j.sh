#!/bin/bash
# GPU call J: per-kernel times at 2,000 streams (ncu launch list of the LM branch and of the codec branch).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/profile_steps.py --workload lm --eager --steps 2 --streams 2000 --kv int8 > gpurun_out/j_plain_lm.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/j_launches_lm2000.csv \
    python tools/profile_steps.py --workload lm --eager --steps 2 --streams 2000 --kv int8 > gpurun_out/j_ncu_lm.log 2>&1
echo "ncu lm rc=$?"
timeout 300 python tools/profile_steps.py --workload mimi --eager --steps 2 --streams 2000 > gpurun_out/j_plain_mimi.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/j_launches_mimi2000.csv \
    python tools/profile_steps.py --workload mimi --eager --steps 2 --streams 2000 > gpurun_out/j_ncu_mimi.log 2>&1
echo "ncu mimi rc=$?"
